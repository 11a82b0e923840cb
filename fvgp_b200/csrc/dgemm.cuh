// FP64 tensor-core GEMM for the blocked Cholesky / inverse.
//
//   C[m][n] = alpha * sum_k Aop[m][k] * Bop[k][n] + beta * C[m][n]        (C row-major)
//
// Operand layouts (template flags): an operand is either "K-major" (its k index is the
// contiguous one: A[m*lda+k], B[n*ldb+k]) or "MN-major" (A[k*lda+m], B[k*ldb+n]).  The
// three combinations the factorisation needs are
//     <K,K>   SYRK / GEMM / TRSM-by-tile-inverse in POTRF     (C -= A A^T)
//     <K,MN>  the two triangular products of TRTRI            (W = L21 M11, L21 = -M22 W)
//     <MN,MN> LAUUM                                           (P = M^T M)
// Math: mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 -- the only FP64 tensor shape
// sm_100a has; the wider PTX shapes lower to sequences of it, tcgen05 has no f64 kind).
// CTA tile 128x128x16 (or 64x64x16 for small products), 8 warps (2x4) with 64x32 (32x16) warp tiles, 4-stage
// cp.async pipeline.
// Shared-memory rows are padded (+4 doubles) so the 8-byte fragment loads of a
// half-warp hit 16 distinct bank pairs for both layouts.
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace fvgp {

// BM is the algorithmic block of the factorisation (splits, triangular-operand alignment).  The GEMM CTA tile is
// TB x TB with TB = 128 (large products: 1 CTA / SM, 64 accumulators per thread) or TB = 64 (products whose
// 128-tile grid would leave most SMs idle -- the latency-bound diagonal chain of POTRF / TRTRI / LAUUM, where
// a 128 x 128 x 128 product used to run on ONE SM for ~17 us: four times the CTAs, a quarter of the work each,
// 2 CTAs / SM).
constexpr int BM = 128, BN = 128, BK = 16;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 256;
constexpr int KMAJ_STRIDE = BK + 4;                 // 20 doubles
template <int TB> struct GemmTile {
  static constexpr int MNMAJ_STRIDE = TB + 4;                // 132 / 68 doubles
  static constexpr int OPERAND_DOUBLES = TB * KMAJ_STRIDE;   // >= BK * MNMAJ_STRIDE
  static constexpr int STAGE_DOUBLES = 2 * OPERAND_DOUBLES;
  static constexpr size_t SMEM_BYTES = size_t(GEMM_STAGES) * STAGE_DOUBLES * sizeof(double);  // 163840 / 81920
  static constexpr int MFR = TB / 16;   // 8x8 fragments per warp along m (warp grid 2 x 4)
  static constexpr int NFR = TB / 32;   // ... along n
  static constexpr int QN = TB / 32;    // 16-byte copies per thread, operand and stage
};

enum GemmFlags : int {
  GEMM_LOWER = 1,       // square tile grid, tiles strictly above the diagonal are skipped
  GEMM_KB_FROM_M = 2,   // k starts at the tile's first row     (A^T lower-triangular: A[k][m], k >= m)
  GEMM_KB_FROM_N = 4,   // k starts at the tile's first column  (B lower-triangular:   B[k][n], k >= n)
  GEMM_KE_FROM_M = 8,   // k ends at the tile's last row        (A lower-triangular:   A[m][k], k <= m)
};

struct GemmArgs {
  const double* A;
  const double* B;
  double* C;
  int M, N, K;
  long long lda, ldb, ldc;
  double alpha, beta;
  int flags;
  int tiles_m, tiles_n;
  long long bstride;  // BATCHED kernels: blockIdx.y selects a problem, every pointer moves by blockIdx.y * bstride doubles
};

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gmem_src, int src_bytes) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Rasterisation: tiles are walked in bands of 8 tile-rows, column by column inside a band,
// so the ~148 CTAs resident at any time share a handful of A and B panels in L2.
__device__ __forceinline__ void map_tile(int id, int tiles_m, int tiles_n, bool lower, int& tm, int& tn) {
  constexpr int G = 8;
  if (!lower) {
    const int per_group = G * tiles_n;
    const int grp = id / per_group;
    const int r0 = grp * G;
    const int gh = min(G, tiles_m - r0);
    const int l = id - grp * per_group;
    tn = l / gh;
    tm = r0 + l % gh;
    return;
  }
  int b = 0, off = 0, r0 = 0, gh = 0;
  for (;;) {
    r0 = b * G;
    gh = min(G, tiles_m - r0);
    const int cnt = r0 * gh + gh * (gh + 1) / 2;
    if (id < off + cnt) break;
    off += cnt;
    ++b;
  }
  int l = id - off;
  if (l < r0 * gh) {
    tn = l / gh;
    tm = r0 + l % gh;
    return;
  }
  l -= r0 * gh;
  int c = 0;
  while (l >= gh - c) {
    l -= gh - c;
    ++c;
  }
  tn = r0 + c;
  tm = r0 + c + l;
}

// Stage one 128 x 16 operand tile.  `rows` is the operand's extent along m (or n).
template <bool MN_MAJOR, int TB>
__device__ __forceinline__ void load_operand(double* s, const double* __restrict__ g, long long ld, int row0,
                                             int rows, int k0, int kend, int tid) {
#pragma unroll
  for (int q = 0; q < GemmTile<TB>::QN; ++q) {
    const int c = tid + GEMM_THREADS * q;
    if (!MN_MAJOR) {
      const int r = c >> 3, kc = (c & 7) * 2;
      const int gr = row0 + r, gk = k0 + kc;
      int valid = (gr < rows) ? min(max(kend - gk, 0), 2) : 0;
      const double* src = valid ? g + (long long)gr * ld + gk : g;
      cp_async16(s + r * KMAJ_STRIDE + kc, src, valid * 8);
    } else {
      const int kk = c / (TB / 2), mc = (c % (TB / 2)) * 2;
      const int gk = k0 + kk, gm = row0 + mc;
      int valid = (gk < kend) ? min(max(rows - gm, 0), 2) : 0;
      const double* src = valid ? g + (long long)gk * ld + gm : g;
      cp_async16(s + kk * GemmTile<TB>::MNMAJ_STRIDE + mc, src, valid * 8);
    }
  }
}

// BATCHED: gridDim.y independent problems of identical shape whose operands all live at the same offsets of
// equally sized workspace slots (population evaluation, dense_linalg.cu); the plain instantiation is unchanged.
template <bool A_MN, bool B_MN, int TB, bool BATCHED = false>
__global__ void __launch_bounds__(GEMM_THREADS, (TB == 128 ? 1 : 2)) dgemm_mma_kernel(const GemmArgs p) {
  using T = GemmTile<TB>;
  const long long boff = BATCHED ? (long long)blockIdx.y * p.bstride : 0ll;
  const double* gA = p.A + boff;
  const double* gB = p.B + boff;
  double* gC = p.C + boff;
  constexpr int MFR = T::MFR, NFR = T::NFR, QN = T::QN;
  constexpr int MNMAJ_STRIDE = T::MNMAJ_STRIDE, OPERAND_DOUBLES = T::OPERAND_DOUBLES, STAGE_DOUBLES = T::STAGE_DOUBLES;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 2) * (TB / 2), wn0 = (warp & 3) * (TB / 4);

  int tm, tn;
  map_tile(blockIdx.x, p.tiles_m, p.tiles_n, (p.flags & GEMM_LOWER) != 0, tm, tn);
  const int m0 = tm * TB, n0 = tn * TB;

  int kb = 0, ke = p.K;
  if (p.flags & GEMM_KB_FROM_M) kb = max(kb, m0);
  if (p.flags & GEMM_KB_FROM_N) kb = max(kb, n0);
  if (p.flags & GEMM_KE_FROM_M) ke = min(ke, m0 + TB);
  const int kt_total = ke > kb ? (ke - kb + BK - 1) / BK : 0;

  double acc[MFR][NFR][2];
#pragma unroll
  for (int i = 0; i < MFR; ++i)
#pragma unroll
    for (int j = 0; j < NFR; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // Per-thread copy descriptors, computed once: every k-tile except a ragged last one is a pure
  // pointer bump (the generic path costs ~150 integer instructions per k-tile and sat between the
  // barrier and the first DMMA of every warp).
  const double* a_src[QN];
  const double* b_src[QN];
  int a_bytes[QN], b_bytes[QN], a_off[QN], b_off[QN];
#pragma unroll
  for (int q = 0; q < QN; ++q) {
    const int c = tid + GEMM_THREADS * q;
    if (!A_MN) {
      const int r = c >> 3, kc = (c & 7) * 2;
      a_off[q] = r * KMAJ_STRIDE + kc;
      a_bytes[q] = (m0 + r < p.M) ? 16 : 0;
      a_src[q] = a_bytes[q] ? gA + (long long)(m0 + r) * p.lda + kb + kc : gA;
    } else {
      const int kk = c / (TB / 2), mc = (c % (TB / 2)) * 2;
      a_off[q] = kk * MNMAJ_STRIDE + mc;
      a_bytes[q] = 8 * min(max(p.M - (m0 + mc), 0), 2);
      a_src[q] = a_bytes[q] ? gA + (long long)(kb + kk) * p.lda + m0 + mc : gA;
    }
    if (!B_MN) {
      const int r = c >> 3, kc = (c & 7) * 2;
      b_off[q] = r * KMAJ_STRIDE + kc;
      b_bytes[q] = (n0 + r < p.N) ? 16 : 0;
      b_src[q] = b_bytes[q] ? gB + (long long)(n0 + r) * p.ldb + kb + kc : gB;
    } else {
      const int kk = c / (TB / 2), nc = (c % (TB / 2)) * 2;
      b_off[q] = kk * MNMAJ_STRIDE + nc;
      b_bytes[q] = 8 * min(max(p.N - (n0 + nc), 0), 2);
      b_src[q] = b_bytes[q] ? gB + (long long)(kb + kk) * p.ldb + n0 + nc : gB;
    }
  }
  const long long a_step = A_MN ? (long long)BK * p.lda : BK;
  const long long b_step = B_MN ? (long long)BK * p.ldb : BK;

  auto issue = [&](int kt) {
    double* st = smem + (kt % GEMM_STAGES) * STAGE_DOUBLES;
    const int k0 = kb + kt * BK;
    if (k0 + BK <= ke) {
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        cp_async16(st + a_off[q], a_src[q] + (a_bytes[q] ? kt * a_step : 0), a_bytes[q]);
        cp_async16(st + OPERAND_DOUBLES + b_off[q], b_src[q] + (b_bytes[q] ? kt * b_step : 0), b_bytes[q]);
      }
    } else {
      load_operand<A_MN, TB>(st, gA, p.lda, m0, p.M, k0, ke, tid);
      load_operand<B_MN, TB>(st + OPERAND_DOUBLES, gB, p.ldb, n0, p.N, k0, ke, tid);
    }
  };

#pragma unroll
  for (int s = 0; s < GEMM_STAGES - 1; ++s) {
    if (s < kt_total) issue(s);
    cp_async_commit();
  }

  for (int kt = 0; kt < kt_total; ++kt) {
    cp_async_wait<GEMM_STAGES - 2>();
    __syncthreads();

    const double* As = smem + (kt % GEMM_STAGES) * STAGE_DOUBLES;
    const double* Bs = As + OPERAND_DOUBLES;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double a[MFR], b[NFR];
#pragma unroll
      for (int i = 0; i < MFR; ++i)
        a[i] = A_MN ? As[(4 * kk + t) * MNMAJ_STRIDE + wm0 + 8 * i + g] : As[(wm0 + 8 * i + g) * KMAJ_STRIDE + 4 * kk + t];
#pragma unroll
      for (int j = 0; j < NFR; ++j)
        b[j] = B_MN ? Bs[(4 * kk + t) * MNMAJ_STRIDE + wn0 + 8 * j + g] : Bs[(wn0 + 8 * j + g) * KMAJ_STRIDE + 4 * kk + t];
#pragma unroll
      for (int i = 0; i < MFR; ++i)
#pragma unroll
        for (int j = 0; j < NFR; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      if (kk == 0) {
        // refill the stage freed by the barrier above only after this warp has fed the tensor pipe
        if (kt + GEMM_STAGES - 1 < kt_total) issue(kt + GEMM_STAGES - 1);
        cp_async_commit();
      }
    }
  }
  cp_async_wait<0>();

  // Epilogue: each lane owns two adjacent columns -> 16-byte accesses, 64 B per row per quad.
  const double alpha = p.alpha, beta = p.beta;
#pragma unroll
  for (int i = 0; i < MFR; ++i) {
    const int row = m0 + wm0 + 8 * i + g;
    if (row >= p.M) continue;
    double* crow = gC + (long long)row * p.ldc;
#pragma unroll
    for (int j = 0; j < NFR; ++j) {
      const int col = n0 + wn0 + 8 * j + 2 * t;
      if (col + 1 < p.N) {
        double2 v = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
        if (beta != 0.0) {
          const double2 old = *reinterpret_cast<const double2*>(crow + col);
          v.x += beta * old.x;
          v.y += beta * old.y;
        }
        *reinterpret_cast<double2*>(crow + col) = v;
      } else if (col < p.N) {
        double v = alpha * acc[i][j][0];
        if (beta != 0.0) v += beta * crow[col];
        crow[col] = v;
      }
    }
  }
}

// Host launcher.  Requirements (checked): even leading dimensions and 16-byte aligned
// bases (cp.async 16 B and double2 epilogue).
// FVGP_GEMM_SMALL_TILES=0 disables the 64 x 64 tile variant (A/B on the GPU box).
inline bool gemm_small_tiles_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FVGP_GEMM_SMALL_TILES");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

template <bool A_MN, bool B_MN, int TB, bool BATCHED = false>
inline int launch_gemm_tile(cudaStream_t st, GemmArgs p, int batch = 1) {
  static bool configured = false;
  if (!configured) {
    FVGP_CUDA_OK(cudaFuncSetAttribute(dgemm_mma_kernel<A_MN, B_MN, TB, BATCHED>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmTile<TB>::SMEM_BYTES));
    configured = true;
  }
  p.tiles_m = (p.M + TB - 1) / TB;
  p.tiles_n = (p.N + TB - 1) / TB;
  long long tiles;
  if (p.flags & GEMM_LOWER) {
    FVGP_REQUIRE(p.tiles_m == p.tiles_n);
    tiles = (long long)p.tiles_m * (p.tiles_m + 1) / 2;
  } else {
    tiles = (long long)p.tiles_m * p.tiles_n;
  }
  launch(dgemm_mma_kernel<A_MN, B_MN, TB, BATCHED>, dim3((unsigned)tiles, (unsigned)batch), GEMM_THREADS,
         GemmTile<TB>::SMEM_BYTES, st, p);
  FVGP_LAUNCH_OK();
  return 0;
}

// Host launcher.  Requirements (checked): even leading dimensions and 16-byte aligned
// bases (cp.async 16 B and double2 epilogue).  batch > 1: the same product for `batch` problems whose operands sit
// `bstride` doubles apart (bstride even).
template <bool A_MN, bool B_MN>
inline int launch_gemm(cudaStream_t st, const double* A, long long lda, const double* B, long long ldb, double* C,
                       long long ldc, int M, int N, int K, double alpha, double beta, int flags, int batch = 1,
                       long long bstride = 0) {
  if (M <= 0 || N <= 0) return 0;
  FVGP_REQUIRE((lda % 2 == 0) && (ldb % 2 == 0) && (ldc % 2 == 0));
  FVGP_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0));
  FVGP_REQUIRE(batch >= 1 && batch <= 65535 && (batch == 1 || bstride % 2 == 0));
  GemmArgs p;
  p.A = A, p.B = B, p.C = C, p.M = M, p.N = N, p.K = K;
  p.lda = lda, p.ldb = ldb, p.ldc = ldc, p.alpha = alpha, p.beta = beta, p.flags = flags;
  p.tiles_m = p.tiles_n = 0;
  p.bstride = batch > 1 ? bstride : 0;
  // 128-tiles when they fill at least half the SMs; otherwise the problem is latency bound and more, smaller CTAs win
  const long long tm = (M + BM - 1) / BM, tn = (N + BN - 1) / BN;
  const long long tiles128 = ((flags & GEMM_LOWER) ? tm * (tm + 1) / 2 : tm * tn) * batch;
  // In-place products (C aliases an operand: the TRSM / LAUUM leaves, N <= 128) rely on ONE CTA owning complete rows
  // of the output -- it has consumed its operand rows before the epilogue writes them; 64-wide tiles would let a
  // neighbouring CTA overwrite columns that are still being read.
  const bool in_place = (const double*)C == A || (const double*)C == B;
  const bool small = !in_place && tiles128 * 2 < sm_count() && gemm_small_tiles_enabled();
  if (batch > 1) {
    if (small) return launch_gemm_tile<A_MN, B_MN, 64, true>(st, p, batch);
    return launch_gemm_tile<A_MN, B_MN, 128, true>(st, p, batch);
  }
  if (small) return launch_gemm_tile<A_MN, B_MN, 64>(st, p);
  return launch_gemm_tile<A_MN, B_MN, 128>(st, p);
}

}  // namespace fvgp
