// FP64-accurate GEMM on the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators) -- the one route past the
// DMMA roof of 37 TFLOP/s for the trailing updates of the dense factorisation (gp_lin_alg.py:237-269 calls LAPACK
// dpotrf, whose flops are these updates).  Ozaki-type error-free splitting (Ozaki, Ogita, Oishi, Rump 2012; INT8
// variant: Ootomo, Ozaki, Yokota 2024):
//
//   C -= A B^T,  A (m x k), B (n x k) FP64, rows contiguous.
//   1. every ROW of A (and of B) is scaled by a power of two 2^-e_i so that |a'| < 1, and cut into S slices of 6 bits
//      (round-to-nearest digits q_s in [-64, 64], remainders exact in FP64):  a' = sum_s q_s 2^(-6 s) + O(2^(-6 S - 1));
//   2. integer products of slices are EXACT in int32 (|q q'| <= 2^12, K <= 2^19), so
//        A B^T = 2^(e_i + f_j) sum_g 2^(-6 g) G_g,   G_g = sum_{s+t=g} A_s B_t^T,   g = 2 .. S + 1
//      (terms with s + t > S + 1 are below the truncation error and dropped).  The slices of A are stored side by
//      side ([A_1 | A_2 | ... | A_S], one int8 row of S k bytes per matrix row) and those of B in REVERSE order, which
//      makes every G_g ONE int8 GEMM with K = (g - 1) k on contiguous operands;
//   3. every G_g leaves its GEMM as int32 (the sm100 epilogue builder has no int32 -> f64 store path); ONE combine pass
//      per column block converts them to FP64 (exact), adds 2^(-6 g) G_g smallest terms first, applies the row /
//      column exponents and updates C (S int32 reads + one FP64 read-modify-write per entry: ~1/4 of the MMA time).
//   S (S + 1) / 2 integer MACs per FP64 MAC: S = 8 -> 36, error ~2^-47 relative to the row maxima.
//
// The int8 GEMM is a CuTe / CUTLASS sm100 collective (TMA loads, tcgen05.mma kind::i8 issued by one thread, int32
// accumulators in TMEM, tcgen05.ld epilogue) instantiated inside this file; SASS: UTCIMMA, LDTM, UTMALDG.
#include "../../include/fvgp_b200.h"
#include "common.cuh"

#ifdef FVGP_HAVE_CUTLASS
#include <cstdint>
#include <cstdlib>
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"

namespace fvgp {
namespace oz {
using namespace cute;

using LayoutA = cutlass::layout::RowMajor;     // A slices: M x K, K contiguous
using LayoutB = cutlass::layout::ColumnMajor;  // B slices: N rows of K contiguous bytes = K x N column-major
using LayoutC = cutlass::layout::RowMajor;
// Tile configurations of the same collective (FVGP_OZAKI_TILE): 1 = one SM per 128 x 128 x 128 tile; 2 = a CTA pair
// (cta_group::2, cluster 2 x 1; SASS UTCIMMA.2CTA) on 256 x 128 x 128; 3 (default) = a CTA pair on 256 x 256 x 128;
// 4 = tile 2 in a 2 x 2 cluster.  Raw GEMM rate (tools/i8_rate_probe.py, profiles/r02/i8_rate_probe.v17.log; nominal
// dense INT8 peak 2.25 PMAC/s): 32768 x 4096 x 16384: 1.42 / 1.71 / 2.00 / 1.56 PMAC/s; the POTRI shape
// 6272 x 24912 x 100352: 1.38 / 1.23 / 2.03 / 1.47 -- the wide tile halves the operand traffic per MAC, which is what
// bounds the narrower ones once the operands stop fitting L2.
static unsigned long long g_i8_macs = 0;  // int8 multiply-accumulates launched so far (m n K per GEMM): bench.py's roofline

template <class MmaTileShape, class ClusterShape>
struct I8Gemm {
  // D (int32) = acc
  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
      cutlass::epilogue::collective::EpilogueTileAuto, int32_t, int32_t, int32_t, LayoutC, 4, int32_t, LayoutC, 4,
      cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, int8_t, LayoutA, 16, int8_t, LayoutB, 16, int32_t, MmaTileShape,
      ClusterShape,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

  static int run(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* D, int64_t ldd, int m, int n, int K,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
    auto sa = cute::make_stride(lda, Int<1>{}, int64_t(0));
    auto sb = cute::make_stride(ldb, Int<1>{}, int64_t(0));
    auto sc = cute::make_stride(ldd, Int<1>{}, int64_t(0));
    typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {m, n, K, 1}, {A, sa, B, sb},
                                  {{1, 0}, D, sc, D, sc}};
    Gemm gemm;
    if (gemm.can_implement(args) != cutlass::Status::kSuccess) {
      fprintf(stderr, "[fvgp_b200] ozaki: int8 GEMM %d x %d x %d cannot be implemented (alignment?)\n", m, n, K);
      return FVGP_ERR_ARG;
    }
    if (Gemm::get_workspace_size(args) > ws_bytes) return FVGP_ERR_ARG;
    if (gemm.initialize(args, ws, st) != cutlass::Status::kSuccess) return FVGP_ERR_CUDA;
    if (gemm.run(st) != cutlass::Status::kSuccess) return FVGP_ERR_CUDA;
    __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
    __atomic_fetch_add(&g_i8_macs, (unsigned long long)m * (unsigned long long)n * (unsigned long long)K, __ATOMIC_RELAXED);
    return 0;
  }
};

using I8Gemm1Sm = I8Gemm<Shape<_128, _128, _128>, Shape<_1, _1, _1>>;
using I8Gemm2Sm = I8Gemm<Shape<_256, _128, _128>, Shape<_2, _1, _1>>;
using I8Gemm2SmWide = I8Gemm<Shape<_256, _256, _128>, Shape<_2, _1, _1>>;  // FVGP_OZAKI_TILE=3
using I8Gemm2SmCl4 = I8Gemm<Shape<_256, _128, _128>, Shape<_2, _2, _1>>;   // FVGP_OZAKI_TILE=4

static int g_tile = -1;

// D (m x n int32, ldd) = A (m x K int8, lda) B^T (n x K int8, ldb)
static int i8_gemm(const int8_t* A, int64_t lda, const int8_t* B, int64_t ldb, int32_t* D, int64_t ldd, int m, int n, int K,
                   void* ws, size_t ws_bytes, cudaStream_t st) {
  if (g_tile < 0) {
    const char* e = getenv("FVGP_OZAKI_TILE");
    const int v = e ? atoi(e) : 3;
    g_tile = (v >= 1 && v <= 4) ? v : 3;
  }
  if (g_tile == 2) return I8Gemm2Sm::run(A, lda, B, ldb, D, ldd, m, n, K, ws, ws_bytes, st);
  if (g_tile == 3) return I8Gemm2SmWide::run(A, lda, B, ldb, D, ldd, m, n, K, ws, ws_bytes, st);
  if (g_tile == 4) return I8Gemm2SmCl4::run(A, lda, B, ldb, D, ldd, m, n, K, ws, ws_bytes, st);
  return I8Gemm1Sm::run(A, lda, B, ldb, D, ldd, m, n, K, ws, ws_bytes, st);
}

constexpr int OZ_BITS = 6;
constexpr int OZ_MAX_SLICES = 10;

// One warp per row: row maximum -> exponent e (|a| 2^-e < 1), then S round-to-nearest 6-bit digits per entry, written
// into the forward ([A_1 | ... | A_S]) and / or reversed ([A_S | ... | A_1]) slice rows.  k % 4 == 0.
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ A, long long lda, long long m, int k, int S,
                                                       int8_t* fwd, int8_t* rev, int* exps) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= m) return;
  const double* a = A + row * lda;
  // rows are 16-byte aligned (lda even, column offsets even) and k % 16 == 0: 16 entries per lane and trip, one 16-byte
  // store per slice (the 4-entry version ran at half the HBM rate: profiles/r02/launches_bench_n50k.v21.txt)
  const bool vec = ((reinterpret_cast<uintptr_t>(a) & 15) == 0) && (k % 16 == 0);
  double mx = 0.0;
  if (vec) {
    for (int c = lane * 2; c < k; c += 64) {
      const double2 t = *reinterpret_cast<const double2*>(a + c);
      mx = fmax(mx, fmax(fabs(t.x), fabs(t.y)));
    }
  } else {
    for (int c = lane; c < k; c += 32) mx = fmax(mx, fabs(a[c]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0 && isfinite(mx)) e = ilogb(mx) + 1;  // 2^(e-1) <= mx < 2^e
  if (lane == 0) exps[row] = e;
  const long long rowbytes = (long long)S * k;
  int8_t* f = fwd != nullptr ? fwd + row * rowbytes : nullptr;
  int8_t* r = rev != nullptr ? rev + row * rowbytes : nullptr;
  if (vec) {
    for (int c0 = lane * 16; c0 < k; c0 += 512) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; u += 2) {
        const double2 t = *reinterpret_cast<const double2*>(a + c0 + u);
        v[u] = scalbn(t.x, -e);
        v[u + 1] = scalbn(t.y, -e);
      }
      for (int s = 0; s < S; ++s) {
        int4 q;
        signed char* qq = reinterpret_cast<signed char*>(&q);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const double t = v[u] * 64.0;
          const double d = rint(t);
          v[u] = t - d;  // exact: |t| < 2^7 with at most 53 significant bits, d its nearest integer
          qq[u] = (signed char)(int)d;
        }
        if (f != nullptr) *reinterpret_cast<int4*>(f + (long long)s * k + c0) = q;
        if (r != nullptr) *reinterpret_cast<int4*>(r + (long long)(S - 1 - s) * k + c0) = q;
      }
    }
    return;
  }
  for (int c0 = lane * 4; c0 < k; c0 += 128) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = scalbn(a[c0 + u], -e);
    for (int s = 0; s < S; ++s) {
      char4 q;
      signed char* qq = reinterpret_cast<signed char*>(&q);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double t = v[u] * 64.0;
        const double d = rint(t);
        v[u] = t - d;  // exact: |t| < 2^7 with at most 53 significant bits, d its nearest integer
        qq[u] = (signed char)(int)d;
      }
      if (f != nullptr) *reinterpret_cast<char4*>(f + (long long)s * k + c0) = q;
      if (r != nullptr) *reinterpret_cast<char4*>(r + (long long)(S - 1 - s) * k + c0) = q;
    }
  }
}

// C[i][j] += sign * 2^(ea[i] + eb[j]) * sum_g 2^(-6 g) G_g[i][j], g = S + 1 (first plane) .. 2 (last plane); planes are
// `plane` int32 apart.  lower: only entries with j <= i + diag (block-local form of column <= row).
__global__ void __launch_bounds__(256) oz_combine_kernel(double* __restrict__ C, long long ldc, const int32_t* __restrict__ G,
                                                         long long ldg, long long plane, int S, long long m, long long n,
                                                         const int* __restrict__ ea, const int* __restrict__ eb,
                                                         double sign, int lower, long long diag) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i0 = (long long)blockIdx.y * 8;
  if (j >= n) return;
  const int ej = eb[j];
  for (long long i = i0; i < min(m, i0 + 8); ++i) {
    if (lower && j > i + diag) continue;
    double acc = 0.0;
    double w = scalbn(1.0, -OZ_BITS * (S + 1));
    for (int q = 0; q < S; ++q) {  // plane q holds g = S + 1 - q
      acc = fma((double)G[(long long)q * plane + i * ldg + j], w, acc);
      w *= 64.0;
    }
    C[i * ldc + j] += sign * scalbn(acc, ea[i] + ej);
  }
}

}  // namespace oz
}  // namespace fvgp

using namespace fvgp;

extern "C" {

int fvgp_ozaki_available(void) { return 1; }

#define FVGP_OZAKI_PARTIAL (-100)  /* a GEMM failed after part of C had been updated: C is inconsistent */

// scratch bytes for C (m x n) -= A (m x k) B^T (n x k) with S slices, processed in column blocks of `nblock`
int64_t fvgp_ozaki_work_bytes(int64_t m, int64_t n, int64_t k, int slices, int64_t nblock) {
  const int64_t nb = nblock < n ? nblock : n;
  const int64_t bytes_a = (int64_t)slices * k * m, bytes_b = (int64_t)slices * k * n;
  return bytes_a + bytes_b + 2 * (m + n) * (int64_t)sizeof(int) + (int64_t)slices * m * ((nb + 3) / 4 * 4) * (int64_t)sizeof(int32_t) +
         (8 << 20) + 8192;
}

// C (m x n, ldc) += sign * A (m x k, lda) * B (n x k, ldb)^T, all FP64 row-major.  lower != 0: only entries (i, j) with
// j <= i + diag are updated (SYRK / triangular targets; diag = global row of C's first row minus global column of its
// first column).  same_ab != 0: B is A (n == m): the panel is sliced once.
int fvgp_ozaki_gemm_nt(double* d_C, int64_t ldc, const double* d_A, int64_t lda, const double* d_B, int64_t ldb, int64_t m,
                       int64_t n, int64_t k, double sign, int lower, int64_t diag, int same_ab, int slices, int64_t nblock,
                       void* d_work, int64_t work_bytes, void* stream) {
  using namespace fvgp::oz;
  FVGP_REQUIRE(m > 0 && n > 0 && k > 0 && k % 16 == 0 && slices >= 2 && slices <= OZ_MAX_SLICES && nblock >= 128);
  FVGP_REQUIRE(work_bytes >= fvgp_ozaki_work_bytes(m, n, k, slices, nblock));
  FVGP_REQUIRE(!same_ab || (m == n));
  cudaStream_t st = (cudaStream_t)stream;
  const int S = slices;
  const int64_t rowbytes = (int64_t)S * k;
  char* w = (char*)d_work;
  auto take = [&](int64_t bytes) {
    char* p = w;
    w += (bytes + 255) / 256 * 256;
    return p;
  };
  int8_t* a_fwd = (int8_t*)take(rowbytes * m);
  int8_t* b_rev = (int8_t*)take(rowbytes * n);
  int* ea = (int*)take(m * sizeof(int));
  int* eb = (int*)take(n * sizeof(int));
  const int64_t nb = nblock < n ? nblock : n;
  const int64_t ldg = (nb + 3) / 4 * 4;
  const int64_t plane = m * ldg;
  int32_t* G = (int32_t*)take((int64_t)S * plane * sizeof(int32_t));
  void* ws = take(8 << 20);
  if (same_ab) {
    launch(oz_slice_kernel, (unsigned)((m * 32 + 255) / 256), 256, 0, st, d_A, (long long)lda, (long long)m, (int)k, S, a_fwd,
           b_rev, ea);
    eb = ea;
  } else {
    launch(oz_slice_kernel, (unsigned)((m * 32 + 255) / 256), 256, 0, st, d_A, (long long)lda, (long long)m, (int)k, S, a_fwd,
           (int8_t*)nullptr, ea);
    launch(oz_slice_kernel, (unsigned)((n * 32 + 255) / 256), 256, 0, st, d_B, (long long)ldb, (long long)n, (int)k, S,
           (int8_t*)nullptr, b_rev, eb);
  }
  FVGP_LAUNCH_OK();
  bool touched = false;  // has any entry of C been updated yet?  (a refusal before that leaves C intact)
  for (int64_t j0 = 0; j0 < n; j0 += nb) {
    const int64_t nj = (n - j0) < nb ? (n - j0) : nb;
    // rows that can touch this column block (lower: i + diag >= j0)
    int64_t i0 = 0;
    if (lower) {
      i0 = j0 - diag;
      if (i0 < 0) i0 = 0;
      i0 = i0 / 128 * 128;
      if (i0 >= m) break;
    }
    const int64_t mi = m - i0;
    for (int g = S + 1; g >= 2; --g) {  // plane S + 1 - g
      const int K = (int)((g - 1) * k);
      int rc = i8_gemm(a_fwd + i0 * rowbytes, rowbytes, b_rev + j0 * rowbytes + (int64_t)(S - g + 1) * k, rowbytes,
                       G + (int64_t)(S + 1 - g) * plane, ldg, (int)mi, (int)nj, K, ws, 8 << 20, st);
      if (rc != 0) return touched ? FVGP_OZAKI_PARTIAL : rc;
    }
    touched = true;
    dim3 grid((unsigned)((nj + 255) / 256), (unsigned)((mi + 7) / 8));
    launch(oz_combine_kernel, grid, 256, 0, st, d_C + i0 * ldc + j0, (long long)ldc, (const int32_t*)G, (long long)ldg,
           (long long)plane, S, (long long)mi, (long long)nj, (const int*)(ea + i0), (const int*)(eb + j0), sign, lower,
           (long long)(diag + i0 - j0));
    FVGP_LAUNCH_OK();
  }
  return 0;
}

unsigned long long fvgp_ozaki_mac_count(void) { return fvgp::oz::g_i8_macs; }

// Measurement hook (tools/i8_rate_probe.py): seconds per launch of the raw int8 GEMM m x n x K (operands filled with
// a fixed byte pattern, int32 output) for tile configuration `tile` (1..4, see i8_gemm), best of `reps`; < 0 on error.
double fvgp_ozaki_i8_seconds(int64_t m, int64_t n, int64_t K, int tile, int reps, void* stream) {
  using namespace fvgp::oz;
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 0 || n <= 0 || K <= 0 || K % 16 != 0 || tile < 1 || tile > 4) return -1.0;
  int8_t *A = nullptr, *B = nullptr;
  int32_t* D = nullptr;
  void* ws = nullptr;
  const int64_t ldd = (n + 3) / 4 * 4;
  if (cudaMalloc(&A, m * K) != cudaSuccess || cudaMalloc(&B, n * K) != cudaSuccess ||
      cudaMalloc(&D, m * ldd * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&ws, 8 << 20) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(A), cudaFree(B), cudaFree(D), cudaFree(ws);
    return -2.0;
  }
  cudaMemsetAsync(A, 3, m * K, st);
  cudaMemsetAsync(B, 5, n * K, st);
  const int old = g_tile;
  g_tile = tile;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  double best = 1e30;
  int rc = 0;
  for (int r = 0; r < reps + 1 && rc == 0; ++r) {  // first launch = warm-up
    cudaEventRecord(e0, st);
    rc = i8_gemm(A, K, B, K, D, ldd, (int)m, (int)n, (int)K, ws, 8 << 20, st);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms * 1e-3 < best) best = ms * 1e-3;
  }
  g_tile = old;
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(A), cudaFree(B), cudaFree(D), cudaFree(ws);
  return rc == 0 ? best : -3.0;
}

}  // extern "C"

#else  // built without the CUTLASS headers: the entry points exist and report that the path is unavailable

extern "C" {
int fvgp_ozaki_available(void) { return 0; }
int64_t fvgp_ozaki_work_bytes(int64_t, int64_t, int64_t, int, int64_t) { return 0; }
int fvgp_ozaki_gemm_nt(double*, int64_t, const double*, int64_t, const double*, int64_t, int64_t, int64_t, int64_t, double,
                       int, int64_t, int, int, int64_t, void*, int64_t, void*) {
  fprintf(stderr, "[fvgp_b200] built without CUTLASS headers: the INT8-slice GEMM is not available\n");
  return FVGP_ERR_ARG;
}
double fvgp_ozaki_i8_seconds(int64_t, int64_t, int64_t, int, int, void*) { return -1.0; }
unsigned long long fvgp_ozaki_mac_count(void) { return 0; }
}
#endif
