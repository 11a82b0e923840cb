// Dense FP64 factorisation suite: recursive blocked Cholesky (POTRF), triangular solves
// (POTRS), log-determinant, and the inverse from the factor (POTRI = TRTRI + LAUUM).
//
// Reference behaviour being replaced (lbl-camera/fvGP):
//   calculate_Chol_factor  gp_lin_alg.py:237-269   scipy cho_factor(lower=True)
//   calculate_Chol_solve   gp_lin_alg.py:289-328   cho_solve
//   calculate_Chol_logdet  gp_lin_alg.py:331-360   2*sum(log|diag|)
//   calculate_inv_from_chol gp_lin_alg.py:1558, and the KV^-1 that the gradient's trace
//   term needs (gp_marginal_likelihood.py:273-274 obtains it through batched LU solves).
//
// Design: everything above 64x64 is recursion over ONE tensor-core GEMM kernel
// (dgemm.cuh, DMMA.8x8x4), so >98% of the N^3/3 (+2N^3/3) flops run at GEMM speed with
// the large inner dimensions a cache-oblivious recursion gives (K up to N/2).  Only the
// 64x64 diagonal tiles are factored by a warp-cooperative shared-memory kernel, which
// also emits each tile's inverse; triangular solves against a tile are then GEMMs with
// that inverse (no scalar TRSM anywhere).  All operations are in place on the lower
// triangle of the row-major matrix.
#include "../../include/fvgp_b200.h"
#include "dgemm.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace fvgp {

unsigned long long g_launches = 0;

constexpr int TS = 128;  // diagonal tile handled by one CTA (Cholesky + inverse of the factor)
constexpr int VS = 64;   // block width of the single-right-hand-side solves
constexpr int VBLK = 2048;  // columns per block of the blocked single-right-hand-side solves
constexpr int TP = TS + 1;
constexpr int HP = TS / 2 + 1;
constexpr size_t POTRF_TILE_SMEM = (size_t(TS) * TP + 2 * (TS / 2) * HP + 2 * TS) * sizeof(double);  // ~198 KB

// ----------------------------------------------------------------------------------------------
// 128x128 diagonal tile, one CTA of 512 threads, everything in shared memory:
//   1. right-looking Cholesky, 4 threads per row in the trailing update;
//   2. inverses of the two 64x64 diagonal blocks of the factor by forward substitution
//      (column per 4 lanes, both blocks concurrently);
//   3. the off-diagonal block of the inverse, -D^-1 C A^-1, staged in the unused upper-right
//      quadrant of the tile.
// Emits L (lower, explicit zeros above the diagonal) in place and the full 128x128 inverse of L,
// so that every triangular solve against this tile is ONE tensor-core GEMM with K = 128.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) potrf_tile_kernel(double* A, long long ld, int nt, double* dinv,
                                                         int* info, int global_row0, long long bstride) {
  extern __shared__ __align__(16) double tile_smem[];
  A += blockIdx.x * bstride, dinv += blockIdx.x * bstride, info += blockIdx.x;  // one CTA per problem of a batch
  double(*T)[TP] = reinterpret_cast<double(*)[TP]>(tile_smem);
  double(*InvA)[HP] = reinterpret_cast<double(*)[HP]>(tile_smem + TS * TP);
  double(*InvD)[HP] = reinterpret_cast<double(*)[HP]>(tile_smem + TS * TP + (TS / 2) * HP);
  double* Dg = tile_smem + TS * TP + 2 * (TS / 2) * HP;
  double* RDg = Dg + TS;  // reciprocals of the diagonal
  constexpr int H = TS / 2;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TS * TS; idx += 512) {
    const int r = idx / TS, c = idx % TS;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nt && c <= r) v = A[(long long)r * ld + c];
    T[r][c] = v;
  }
  __syncthreads();
  const int ri = tid >> 2, part = tid & 3;
  for (int j = 0; j < TS; ++j) {
    const double d = T[j][j];
    if (!(d > 0.0) && tid == 0) atomicCAS(info, 0, global_row0 + j + 1);
    // 1/sqrt(d) by MUFU.RSQ64H + two Newton steps (the IEEE sqrt + divide pair is a ~500-cycle
    // dependent chain in every one of the 128 column steps); column scaling becomes a multiply.
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const double e = fma(-(d * y), y, 1.0);
      y = fma(0.5 * y, e, y);
    }
    double s = d * y;
    s = fma(fma(-s, s, d), 0.5 * y, s);
    if (tid == j) Dg[j] = s, RDg[j] = y;
    if (tid > j && tid < TS) T[tid][j] *= y;
    __syncthreads();
    if (ri > j) {
      const double lij = T[ri][j];
      int k = j + 1 + part;
      for (; k + 12 <= ri; k += 16) {  // loads first, then FMAs, then stores: no load-after-store serialisation
        const double a0 = T[k][j], a1 = T[k + 4][j], a2 = T[k + 8][j], a3 = T[k + 12][j];
        const double c0 = T[ri][k], c1 = T[ri][k + 4], c2 = T[ri][k + 8], c3 = T[ri][k + 12];
        T[ri][k] = fma(-lij, a0, c0);
        T[ri][k + 4] = fma(-lij, a1, c1);
        T[ri][k + 8] = fma(-lij, a2, c2);
        T[ri][k + 12] = fma(-lij, a3, c3);
      }
      for (; k <= ri; k += 4) T[ri][k] = fma(-lij, T[k][j], T[ri][k]);
    }
    __syncthreads();
  }
  {  // inverses of the diagonal blocks: column (tid>>2) of block (tid>>8)
    const int c = (tid >> 2) & (H - 1), base = (tid >> 8) * H;
    double(*Inv)[HP] = (tid >> 8) ? InvD : InvA;
    for (int i = 0; i < H; ++i) {
      double acc0 = 0.0, acc1 = 0.0;
      if (i > c) {
        int k = c + part;
        for (; k + 4 < i; k += 8) {
          acc0 = fma(T[base + i][base + k], Inv[k][c], acc0);
          acc1 = fma(T[base + i][base + k + 4], Inv[k + 4][c], acc1);
        }
        for (; k < i; k += 4) acc0 = fma(T[base + i][base + k], Inv[k][c], acc0);
      }
      double acc = acc0 + acc1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) Inv[i][c] = (i < c) ? 0.0 : (((i == c) ? 1.0 : 0.0) - acc) * RDg[base + i];
      __syncwarp();
    }
  }
  __syncthreads();
  {  // E = C A^-1 into T[0..63][64..127]  (C = T[64+i][k])
    const int i = tid >> 3, j0 = (tid & 7) * 8;
    double e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = 0.0;
    for (int k = 0; k < H; ++k) {
      const double cik = T[H + i][k];
#pragma unroll
      for (int u = 0; u < 8; ++u) e[u] = fma(cik, InvA[k][j0 + u], e[u]);   // InvA[k][j] = 0 for k < j
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) T[i][H + j0 + u] = e[u];
  }
  __syncthreads();
  {  // F = -D^-1 E  -> dinv[64+i][j]
    const int i = tid >> 3, j0 = (tid & 7) * 8;
    double f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) f[u] = 0.0;
    for (int k = 0; k <= i; ++k) {
      const double dik = InvD[i][k];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = fma(dik, T[k][H + j0 + u], f[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) dinv[(H + i) * TS + j0 + u] = -f[u];
  }
  for (int idx = tid; idx < TS * TS; idx += 512) {
    const int r = idx / TS, c = idx % TS;
    if (r < H) dinv[idx] = (c < H) ? InvA[r][c] : 0.0;
    else if (c >= H) dinv[idx] = InvD[r - H][c - H];
    if (r < nt && c < nt) A[(long long)r * ld + c] = (c < r) ? T[r][c] : ((c == r) ? Dg[r] : 0.0);
  }
}

// ----------------------------------------------------------------------------------------------
// Second generation of the tile kernel (same contract, same outputs).  The first one spends ~3600 cycles in each of
// its 128 column steps (two CTA barriers, a one-thread-per-row scaling pass and a 4-threads-per-row rank-1 update
// that re-reads the whole trailing tile from shared memory every column): 150-240 us per tile, which is the
// critical path of every evaluation at N ~ 1e3 (8 tiles = half of a 3 ms LML) and of the panel path at large N.
// Here the factorisation is blocked by panels of 16 columns:
//   * inside a panel the columns stay UNSCALED (column j's rank-1 update carries the factor 1/d_j), so one column
//     step is: read the pivot, reciprocal, update the <= 15 remaining panel columns, ONE barrier;
//   * after the 16 columns, one pass scales the panel and also stores it k-major in a 16 x 128 staging array;
//   * the rank-16 update of the trailing tile runs from registers: each thread owns one 4 x 4 micro-tile of the
//     lower triangle (<= 406 micro-tiles), reads its 2 x 4 panel values per k with 16-byte loads from the staging
//     array and touches the trailing tile once per panel instead of once per column.
// The inverse of the factor (steps 2 and 3 of the first kernel) is unchanged.
// ----------------------------------------------------------------------------------------------
constexpr int PW = 16;  // panel width
constexpr size_t POTRF_TILE2_SMEM = POTRF_TILE_SMEM + size_t(PW) * TS * sizeof(double);  // ~214 KB

__global__ void __launch_bounds__(512) potrf_tile2_kernel(double* A, long long ld, int nt, double* dinv, int* info,
                                                          int global_row0, long long bstride) {
  extern __shared__ __align__(16) double tile_smem[];
  A += blockIdx.x * bstride, dinv += blockIdx.x * bstride, info += blockIdx.x;  // one CTA per problem of a batch
  double(*T)[TP] = reinterpret_cast<double(*)[TP]>(tile_smem);
  double(*InvA)[HP] = reinterpret_cast<double(*)[HP]>(tile_smem + TS * TP);
  double(*InvD)[HP] = reinterpret_cast<double(*)[HP]>(tile_smem + TS * TP + (TS / 2) * HP);
  double* Dg = tile_smem + TS * TP + 2 * (TS / 2) * HP;
  double* RDg = Dg + TS;  // reciprocals of the diagonal
  // staging array of the current panel, k-major; offset rounded so that its rows are 16-byte aligned
  constexpr int P_OFF = ((TS * TP + 2 * (TS / 2) * HP + 2 * TS) + 1) / 2 * 2;
  double(*P)[TS] = reinterpret_cast<double(*)[TS]>(tile_smem + P_OFF);
  __shared__ double Pv[PW];  // pivots d_j of the current panel
  constexpr int H = TS / 2;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TS * TS; idx += 512) {
    const int r = idx / TS, c = idx % TS;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nt && c <= r) v = A[(long long)r * ld + c];
    T[r][c] = v;
  }
  // micro-tile of the trailing update owned by this thread: index tid in the row-major enumeration of the lower
  // triangle of a grid of 4 x 4 micro-tiles (the same for every panel; fewer of them are active as panels advance)
  int ma = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
  while ((ma + 1) * (ma + 2) / 2 <= tid) ++ma;
  while (ma * (ma + 1) / 2 > tid) --ma;
  const int mb = tid - ma * (ma + 1) / 2;
  const int row = tid & (TS - 1), cg = tid >> 7;  // column-step / scaling mapping: one row, 4 column groups
  __syncthreads();
  for (int p0 = 0; p0 < TS; p0 += PW) {
    const int pend = p0 + PW;
    for (int j = p0; j < pend; ++j) {
      // Only 1 / d_j sits on the per-column critical path (reciprocal seed + two Newton steps: 4 dependent FMAs); the
      // square root that the scaling pass needs is taken once per panel, four independent chains per thread.
      const double d = T[j][j];
      if (tid == 0) {
        if (!(d > 0.0)) atomicCAS(info, 0, global_row0 + j + 1);
        Pv[j - p0] = d;
      }
      double r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const double e = fma(-d, r, 1.0);
        r = fma(r, e, r);
      }
      if (row > j) {  // T[row][k] -= T[row][j] T[k][j] / d for the panel columns k > j, rows >= k
        const double lij = T[row][j] * r;
#pragma unroll
        for (int q = 0; q < PW / 4; ++q) {
          const int k = j + 1 + cg + 4 * q;
          if (k < pend && k <= row) T[row][k] = fma(-lij, T[k][j], T[row][k]);
        }
      }
      __syncthreads();
    }
    // scale the panel: L[r][j] = T[r][j] / sqrt(d_j); k-major copy for the register-tiled trailing update
#pragma unroll
    for (int q = 0; q < PW / 4; ++q) {
      const int j = p0 + cg + 4 * q;
      if (row >= j) {
        const double d = Pv[j - p0];
        double y;  // 1/sqrt(d) by MUFU.RSQ64H + two Newton steps
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const double e = fma(-(d * y), y, 1.0);
          y = fma(0.5 * y, e, y);
        }
        if (row == j) {
          double sq = d * y;
          sq = fma(fma(-sq, sq, d), 0.5 * y, sq);
          Dg[j] = sq, RDg[j] = y;
        } else {
          const double v = T[row][j] * y;
          T[row][j] = v;
          P[j - p0][row] = v;
        }
      }
    }
    __syncthreads();
    const int off = pend, mt = (TS - off) / 4;
    if (tid < mt * (mt + 1) / 2) {
      const int r0 = off + 4 * ma, c0 = off + 4 * mb;
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
#pragma unroll
      for (int k = 0; k < PW; ++k) {
        const double2 ra = *reinterpret_cast<const double2*>(&P[k][r0]);
        const double2 rb = *reinterpret_cast<const double2*>(&P[k][r0 + 2]);
        const double2 ca = *reinterpret_cast<const double2*>(&P[k][c0]);
        const double2 cb = *reinterpret_cast<const double2*>(&P[k][c0 + 2]);
        const double rv[4] = {ra.x, ra.y, rb.x, rb.y};
        const double cv[4] = {ca.x, ca.y, cb.x, cb.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(rv[i], cv[jj], acc[i][jj]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) T[r0 + i][c0 + jj] -= acc[i][jj];
    }
    __syncthreads();
  }
  const int part = tid & 3;
  {  // inverses of the diagonal blocks: column (tid>>2) of block (tid>>8)
    const int c = (tid >> 2) & (H - 1), base = (tid >> 8) * H;
    double(*Inv)[HP] = (tid >> 8) ? InvD : InvA;
    for (int i = 0; i < H; ++i) {
      double acc0 = 0.0, acc1 = 0.0;
      if (i > c) {
        int k = c + part;
        for (; k + 4 < i; k += 8) {
          acc0 = fma(T[base + i][base + k], Inv[k][c], acc0);
          acc1 = fma(T[base + i][base + k + 4], Inv[k + 4][c], acc1);
        }
        for (; k < i; k += 4) acc0 = fma(T[base + i][base + k], Inv[k][c], acc0);
      }
      double acc = acc0 + acc1;
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) Inv[i][c] = (i < c) ? 0.0 : (((i == c) ? 1.0 : 0.0) - acc) * RDg[base + i];
      __syncwarp();
    }
  }
  __syncthreads();
  // Columns are interleaved over the 8 threads of a row (thread c owns columns c, c + 8, ..., c + 56): for a fixed
  // u the 8 threads read 8 consecutive doubles (the blocked assignment j0 = 8 c made them hit 2 banks: 4-way conflict
  // on every one of the 8 loads per k, ncu l1tex__data_bank_conflicts 92.6k per tile).
  {  // E = C A^-1 into T[0..63][64..127]  (C = T[64+i][k])
    const int i = tid >> 3, jc = tid & 7;
    double e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = 0.0;
    for (int k = 0; k < H; ++k) {
      const double cik = T[H + i][k];
#pragma unroll
      for (int u = 0; u < 8; ++u) e[u] = fma(cik, InvA[k][jc + 8 * u], e[u]);   // InvA[k][j] = 0 for k < j
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) T[i][H + jc + 8 * u] = e[u];
  }
  __syncthreads();
  {  // F = -D^-1 E  -> dinv[64+i][j]
    const int i = tid >> 3, jc = tid & 7;
    double f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) f[u] = 0.0;
    for (int k = 0; k <= i; ++k) {
      const double dik = InvD[i][k];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = fma(dik, T[k][H + jc + 8 * u], f[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) dinv[(H + i) * TS + jc + 8 * u] = -f[u];
  }
  for (int idx = tid; idx < TS * TS; idx += 512) {
    const int r = idx / TS, c = idx % TS;
    if (r < H) dinv[idx] = (c < H) ? InvA[r][c] : 0.0;
    else if (c >= H) dinv[idx] = InvD[r - H][c - H];
    if (r < nt && c < nt) A[(long long)r * ld + c] = (c < r) ? T[r][c] : ((c == r) ? Dg[r] : 0.0);
  }
}

// ----------------------------------------------------------------------------------------------
// Third-generation tile kernel: 32-column panels, the 32 x 32 diagonal block of each panel factored by ONE WARP IN
// REGISTERS.  The second-generation kernel pays one CTA-wide barrier and a dependent pivot chain per COLUMN (128 of
// them, ~720 cycles each: 138 us per tile, ncu r01) -- the critical path of every evaluation below N ~ 16 000 and 3.4x
// slower than cuSOLVER at N = 8192.  Here a tile is 4 panels x 3 barriers:
//   (1) warp 0, lane r = row r of the 32 x 32 diagonal block (32 doubles in registers): right-looking Cholesky, the
//       pivot travels by shuffle, the finished column through a 32-double shared buffer read back as broadcasts; no
//       CTA barrier inside the block;
//   (2) concurrently: warps 1-3 solve the rows BELOW the block against it (one row per thread in registers, forward
//       substitution with broadcast reads of L), warp 0 inverts the block (one row of the inverse per lane) straight
//       into the diagonal position of the 64 x 64 inverses the epilogue needs;
//   (3) all warps: rank-32 update of the trailing tile from the k-major staged panel (4 x 4 register micro-tiles).
// Epilogue: 32 -> 64 -> 128 assembly of the inverse of the factor by -D^-1 C A^-1 products.  Same outputs as before:
// L in place (explicit zeros above the diagonal) and the full 128 x 128 inverse of L.
// ----------------------------------------------------------------------------------------------
constexpr int DB = 32;  // diagonal sub-block / panel width
// T | InvA | InvD | Dg | RDg | P (32 x 96, k-major panel staging; reused as 2 x 32 x 33 scratch by the epilogue) | cb (2 x 32)
constexpr int TILE3_P_OFF = ((TS * TP + 2 * (TS / 2) * HP + 2 * TS) + 1) / 2 * 2;
constexpr int TILE3_CB_OFF = TILE3_P_OFF + DB * (TS - DB);
constexpr size_t POTRF_TILE3_SMEM = size_t(TILE3_CB_OFF + 2 * DB) * sizeof(double);  // ~221 KB

__device__ __forceinline__ double rsqrt_newton(double d) {  // 1 / sqrt(d), d > 0: MUFU.RSQ64H + two Newton steps
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const double e = fma(-(d * y), y, 1.0);
    y = fma(0.5 * y, e, y);
  }
  return y;
}

__global__ void __launch_bounds__(512) potrf_tile3_kernel(double* A, long long ld, int nt, double* dinv, int* info,
                                                          int global_row0, long long bstride) {
  extern __shared__ __align__(16) double tile_smem[];
  A += blockIdx.x * bstride, dinv += blockIdx.x * bstride, info += blockIdx.x;  // one CTA per problem of a batch
  double(*T)[TP] = reinterpret_cast<double(*)[TP]>(tile_smem);
  double(*InvA)[HP] = reinterpret_cast<double(*)[HP]>(tile_smem + TS * TP);
  double(*InvD)[HP] = reinterpret_cast<double(*)[HP]>(tile_smem + TS * TP + (TS / 2) * HP);
  double* Dg = tile_smem + TS * TP + 2 * (TS / 2) * HP;
  double* RDg = Dg + TS;  // reciprocals of the diagonal of L
  double(*P)[TS - DB] = reinterpret_cast<double(*)[TS - DB]>(tile_smem + TILE3_P_OFF);
  double(*cb)[DB] = reinterpret_cast<double(*)[DB]>(tile_smem + TILE3_CB_OFF);
  constexpr int H = TS / 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < TS * TS; idx += 512) {
    const int r = idx / TS, c = idx % TS;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nt && c <= r) v = A[(long long)r * ld + c];
    T[r][c] = v;
  }
  for (int idx = tid; idx < 2 * H * HP; idx += 512) (&InvA[0][0])[idx] = 0.0;  // InvA and InvD are contiguous
  // micro-tile of the trailing update owned by this thread (row-major enumeration of the lower triangle of 4 x 4 tiles)
  int ma = (int)((sqrtf(8.0f * (float)tid + 1.0f) - 1.0f) * 0.5f);
  while ((ma + 1) * (ma + 2) / 2 <= tid) ++ma;
  while (ma * (ma + 1) / 2 > tid) --ma;
  const int mb = tid - ma * (ma + 1) / 2;
  __syncthreads();
#pragma unroll 1
  for (int p0 = 0; p0 < TS; p0 += DB) {
    // ---- (1) diagonal block in registers, one warp
    if (warp == 0) {
      double a[DB];
#pragma unroll
      for (int c = 0; c < DB; ++c) a[c] = (c <= lane) ? T[p0 + lane][p0 + c] : 0.0;
#pragma unroll
      for (int j = 0; j < DB; ++j) {
        const double d = __shfl_sync(0xffffffffu, a[j], j);
        if (lane == 0 && !(d > 0.0)) atomicCAS(info, 0, global_row0 + p0 + j + 1);
        const double y = rsqrt_newton(d);
        double l = 0.0;
        if (lane > j) {
          l = a[j] * y;
        } else if (lane == j) {
          double sq = d * y;
          sq = fma(fma(-sq, sq, d), 0.5 * y, sq);
          l = sq;
          Dg[p0 + j] = sq, RDg[p0 + j] = y;
        }
        a[j] = l;
        cb[j & 1][lane] = l;
        __syncwarp();
#pragma unroll
        for (int k = j + 1; k < DB; ++k)
          if (lane >= k) a[k] = fma(-l, cb[j & 1][k], a[k]);
      }
#pragma unroll
      for (int c = 0; c < DB; ++c)
        if (c <= lane) T[p0 + lane][p0 + c] = a[c];
    }
    __syncthreads();
    // ---- (2) rows below the block (warps 1..3) | inverse of the block (warp 0)
    if (warp == 0) {
      double s[DB];
#pragma unroll
      for (int c = 0; c < DB; ++c) s[c] = (c == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int c = DB - 1; c >= 0; --c) {  // row `lane` of X with X L = I:  x_c = s_c / L_cc ; s_c' -= x_c L_cc'
        const double x = s[c] * RDg[p0 + c];
        s[c] = x;
#pragma unroll
        for (int c2 = 0; c2 < c; ++c2) s[c2] = fma(-x, T[p0 + c][p0 + c2], s[c2]);
      }
      double(*Inv)[HP] = (p0 < H) ? InvA : InvD;
      const int q0 = p0 & (H - 1);
#pragma unroll
      for (int c = 0; c < DB; ++c) Inv[q0 + lane][q0 + c] = s[c];  // zeros above the diagonal included
    } else {
      const int r = p0 + DB + (tid - 32);
      if (tid - 32 < TS - DB && r < TS) {
        double b[DB];
#pragma unroll
        for (int c = 0; c < DB; ++c) b[c] = T[r][p0 + c];
#pragma unroll
        for (int c = 0; c < DB; ++c) {  // x L^T = b:  x_c = b_c / L_cc ; b_c' -= x_c L_c'c
          const double x = b[c] * RDg[p0 + c];
          b[c] = x;
#pragma unroll
          for (int c2 = c + 1; c2 < DB; ++c2) b[c2] = fma(-x, T[p0 + c2][p0 + c], b[c2]);
        }
#pragma unroll
        for (int c = 0; c < DB; ++c) {
          T[r][p0 + c] = b[c];
          P[c][r - (p0 + DB)] = b[c];
        }
      }
    }
    __syncthreads();
    // ---- (3) rank-32 update of the trailing tile
    const int off = p0 + DB, mt = (TS - off) / 4;
    if (tid < mt * (mt + 1) / 2) {
      const int r0 = 4 * ma, c0 = 4 * mb;  // relative to `off`
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
#pragma unroll 8
      for (int k = 0; k < DB; ++k) {
        const double2 ra = *reinterpret_cast<const double2*>(&P[k][r0]);
        const double2 rb = *reinterpret_cast<const double2*>(&P[k][r0 + 2]);
        const double2 ca = *reinterpret_cast<const double2*>(&P[k][c0]);
        const double2 cb2 = *reinterpret_cast<const double2*>(&P[k][c0 + 2]);
        const double rv[4] = {ra.x, ra.y, rb.x, rb.y};
        const double cv[4] = {ca.x, ca.y, cb2.x, cb2.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fma(rv[i], cv[jj], acc[i][jj]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) T[off + r0 + i][off + c0 + jj] -= acc[i][jj];
    }
    __syncthreads();
  }
  // ---- inverse of the factor, 32 -> 64: Inv[32+i][j] = -(Inv_hi C Inv_lo)[i][j] inside both 64 x 64 blocks
  {
    double(*E32)[DB][DB + 1] = reinterpret_cast<double(*)[DB][DB + 1]>(tile_smem + TILE3_P_OFF);  // 2 x 32 x 33
    const int blk = tid >> 8, i = (tid >> 3) & 31, jc = tid & 7, base = blk * H;
    double(*Inv)[HP] = blk ? InvD : InvA;
    double e[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < DB; ++k) {  // E = C Inv_lo   (Inv_lo[k][j] = 0 for k < j)
      const double cik = T[base + DB + i][base + k];
#pragma unroll
      for (int u = 0; u < 4; ++u) e[u] = fma(cik, Inv[k][jc + 8 * u], e[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) E32[blk][i][jc + 8 * u] = e[u];
    __syncthreads();
    double f[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k <= i; ++k) {  // F = Inv_hi E
      const double dik = Inv[DB + i][DB + k];
#pragma unroll
      for (int u = 0; u < 4; ++u) f[u] = fma(dik, E32[blk][k][jc + 8 * u], f[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) Inv[DB + i][jc + 8 * u] = -f[u];
  }
  __syncthreads();
  // ---- 64 -> 128 (as in the second-generation kernel): E = C A^-1 into T[0..63][64..127], F = -D^-1 E -> dinv
  {
    const int i = tid >> 3, jc = tid & 7;
    double e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) e[u] = 0.0;
    for (int k = 0; k < H; ++k) {
      const double cik = T[H + i][k];
#pragma unroll
      for (int u = 0; u < 8; ++u) e[u] = fma(cik, InvA[k][jc + 8 * u], e[u]);  // InvA[k][j] = 0 for k < j
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) T[i][H + jc + 8 * u] = e[u];
  }
  __syncthreads();
  {
    const int i = tid >> 3, jc = tid & 7;
    double f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) f[u] = 0.0;
    for (int k = 0; k <= i; ++k) {
      const double dik = InvD[i][k];
#pragma unroll
      for (int u = 0; u < 8; ++u) f[u] = fma(dik, T[k][H + jc + 8 * u], f[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) dinv[(H + i) * TS + jc + 8 * u] = -f[u];
  }
  for (int idx = tid; idx < TS * TS; idx += 512) {
    const int r = idx / TS, c = idx % TS;
    if (r < H) dinv[idx] = (c < H) ? InvA[r][c] : 0.0;
    else if (c >= H) dinv[idx] = InvD[r - H][c - H];
    if (r < nt && c < nt) A[(long long)r * ld + c] = (c < r) ? T[r][c] : ((c == r) ? Dg[r] : 0.0);
  }
}

// FVGP_POTRF_TILE=2 selects the second-generation tile kernel (A/B on the GPU box).
static bool use_tile_v2() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FVGP_POTRF_TILE");
    v = (e && atoi(e) == 2) ? 1 : 0;
  }
  return v == 1;
}

// FVGP_POTRF_TILE=1 selects the first-generation tile kernel (A/B on the GPU box).
static bool use_tile_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FVGP_POTRF_TILE");
    v = (e && atoi(e) == 1) ? 1 : 0;
  }
  return v == 1;
}

// dst (rows x cols, ldd) <- src (lds)
__global__ void copy2d_kernel(double* dst, long long ldd, const double* src, long long lds, int rows, int cols,
                              long long bstride) {
  dst += blockIdx.z * bstride, src += blockIdx.z * bstride;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * 16;
  if (c >= cols) return;
  for (int r = r0; r < min(r0 + 16, rows); ++r) dst[(long long)r * ldd + c] = src[(long long)r * lds + c];
}

// dst (cols x rows_pad, ldd) <- src (rows x cols, lds)^T, rows [rows, rows_pad) of the source read as zeros.  32 x 32
// tiles through shared memory, both sides coalesced.
// lower != 0: only the lower triangle of the (square) source is read (r >= c), everything above it counts as zero --
// the strict upper triangle of the factor / inverse buffer is never written and holds garbage.
__global__ void __launch_bounds__(256) transpose_pad_kernel(double* __restrict__ dst, long long ldd,
                                                            const double* __restrict__ src, long long lds, int rows, int cols,
                                                            int rows_pad, int lower) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols && (!lower || r >= c)) ? src[(long long)r * lds + c] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + tx;  // dst row = source column
    if (c < cols && r < rows_pad) dst[(long long)c * ldd + r] = tile[tx][ty + 8 * i];
  }
}

// dst (rows x cols, ldd) <- src rows [row0, row0 + rows) of a lower-triangular matrix, columns [0, cols): entries right
// of the diagonal (c > row0 + r) are written as zeros instead of being read.
__global__ void __launch_bounds__(256) copy_lower_rows_kernel(double* __restrict__ dst, long long ldd,
                                                              const double* __restrict__ src, long long lds, int rows, int cols,
                                                              int row0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * 16;
  if (c >= cols) return;
  for (int r = r0; r < min(r0 + 16, rows); ++r) dst[(long long)r * ldd + c] = (c <= row0 + r) ? src[(long long)r * lds + c] : 0.0;
}

// Zero the strict upper part inside every 128x128 diagonal block (the GEMM k-trims assume
// triangular operands carry explicit zeros there).
__global__ void zero_upper_diag_blocks_kernel(double* A, long long ld, int n, long long bstride) {
  A += blockIdx.y * bstride;
  const int b0 = blockIdx.x * BM;
  for (int idx = threadIdx.x; idx < BM * BM; idx += blockDim.x) {
    const int r = idx / BM, c = idx % BM;
    if (c > r && b0 + c < n && b0 + r < n) A[(long long)(b0 + r) * ld + b0 + c] = 0.0;
  }
}

// ----------------------------------------------------------------------------------------------
// Single right-hand-side triangular solves: one launch per 64-wide tile, leaf solve by the
// stored tile inverse fused with the rank-64 update of the remaining vector.
// ----------------------------------------------------------------------------------------------
// The 64x64 inverse block is staged in shared memory with coalesced, independent 16-byte loads (the previous
// version walked it with 16 dependent global loads per thread: ~3 us of pure latency in each of the 782 steps).
constexpr int VSP = VS + 1;  // padded row of the staged inverse block

__device__ __forceinline__ void stage_inverse_block(const double* __restrict__ dinv, double (*S)[VSP], int tid) {
#pragma unroll
  for (int q = 0; q < (VS * VS / 2) / 256; ++q) {  // 2048 double2 / 256 threads
    const int e = tid + 256 * q, r = e / (VS / 2), c2 = (e % (VS / 2)) * 2;
    const double2 v = *reinterpret_cast<const double2*>(dinv + r * TS + c2);
    S[r][c2] = v.x;
    S[r][c2 + 1] = v.y;
  }
}

__global__ void __launch_bounds__(256) fwd_step_kernel(const double* __restrict__ L, long long ld, int n, int j0,
                                                       const double* __restrict__ dinv, double* w, double* z,
                                                       long long bstride) {
  L += blockIdx.y * bstride, dinv += blockIdx.y * bstride, w += blockIdx.y * bstride, z += blockIdx.y * bstride;
  __shared__ double S[VS][VSP];
  __shared__ double yj[VS], zj[VS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nt = min(VS, n - j0);
  stage_inverse_block(dinv, S, tid);
  if (tid < VS) yj[tid] = tid < nt ? w[j0 + tid] : 0.0;
  __syncthreads();
  {
    const int r = tid >> 2, part = tid & 3;
    double acc = 0.0;
#pragma unroll 4
    for (int k = part; k <= r; k += 4) acc = fma(S[r][k], yj[k], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) {
      zj[r] = acc;
      if (blockIdx.x == 0 && r < nt) z[j0 + r] = acc;
    }
  }
  __syncthreads();
  const int rest0 = j0 + VS;
  const double z0 = zj[lane], z1 = zj[lane + 32];
  for (int i = rest0 + blockIdx.x * 64 + warp; i < min(n, rest0 + (int)(blockIdx.x + 1) * 64); i += 8) {
    const double* row = L + (long long)i * ld + j0;
    double acc = row[lane] * z0 + row[lane + 32] * z1;
    acc = warp_sum(acc);
    if (lane == 0) w[i] -= acc;
  }
}

__global__ void __launch_bounds__(256) bwd_step_kernel(const double* __restrict__ L, long long ld, int n, int j0,
                                                       const double* __restrict__ dinv, double* z, double* x,
                                                       int c_begin, long long bstride) {
  L += blockIdx.y * bstride, dinv += blockIdx.y * bstride, z += blockIdx.y * bstride, x += blockIdx.y * bstride;
  __shared__ double S[VS][VSP];
  __shared__ double zj[VS], xj[VS];
  const int tid = threadIdx.x;
  const int nt = min(VS, n - j0);
  stage_inverse_block(dinv, S, tid);
  if (tid < VS) zj[tid] = tid < nt ? z[j0 + tid] : 0.0;
  __syncthreads();
  {
    const int c = tid >> 2, part = tid & 3;
    double acc = 0.0;
#pragma unroll 4
    for (int r = c + part; r < VS; r += 4) acc = fma(S[r][c], zj[r], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) {
      xj[c] = acc;
      if (blockIdx.x == 0 && c < nt) x[j0 + c] = acc;
    }
  }
  __syncthreads();
  const int c = c_begin + blockIdx.x * 256 + tid;
  if (c < j0) {
    const double* col = L + (long long)j0 * ld + c;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int r = 0;
    for (; r + 4 <= nt; r += 4) {  // four independent loads in flight
      a0 = fma(col[(long long)r * ld], xj[r], a0);
      a1 = fma(col[(long long)(r + 1) * ld], xj[r + 1], a1);
      a2 = fma(col[(long long)(r + 2) * ld], xj[r + 2], a2);
      a3 = fma(col[(long long)(r + 3) * ld], xj[r + 3], a3);
    }
    for (; r < nt; ++r) a0 = fma(col[(long long)r * ld], xj[r], a0);
    z[c] -= (a0 + a1) + (a2 + a3);
  }
}

// Both substitutions of one right-hand side in ONE launch for n <= TRSV_FUSED_MAX: a single CTA per problem walks
// the 64-wide tiles itself (stage the inverse block, leaf product, rank-64 update of the remaining vector, barrier).
// At these sizes the step-per-launch version is a chain of 2 n / 64 dependent ~10 us launches (32 of the 66 launches
// of an N = 1000 evaluation, 20 % of a population step); the arithmetic per row / column is the step kernels' own,
// in the same order (bitwise identical results).  w, z: n doubles of scratch each; x: right-hand side in, solution out.
constexpr int TRSV_FUSED_MAX = 2048;
constexpr int TRSV_FUSED_THREADS = 1024;

__device__ __forceinline__ void stage_inverse_block_wide(const double* __restrict__ dinv, double (*S)[VSP], int tid) {
#pragma unroll
  for (int q = 0; q < (VS * VS / 2) / TRSV_FUSED_THREADS; ++q) {  // 2048 double2 / 1024 threads
    const int e = tid + TRSV_FUSED_THREADS * q, r = e / (VS / 2), c2 = (e % (VS / 2)) * 2;
    const double2 v = *reinterpret_cast<const double2*>(dinv + r * TS + c2);
    S[r][c2] = v.x;
    S[r][c2 + 1] = v.y;
  }
}

__global__ void __launch_bounds__(TRSV_FUSED_THREADS) trsv_fused_kernel(const double* __restrict__ L, long long ld, int n,
                                                                        const double* __restrict__ dinv_base, double* w,
                                                                        double* z, double* x, long long bstride) {
  L += blockIdx.x * bstride, dinv_base += blockIdx.x * bstride;
  w += blockIdx.x * bstride, z += blockIdx.x * bstride, x += blockIdx.x * bstride;
  __shared__ double S[VS][VSP];
  __shared__ double yj[VS], zj[VS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = TRSV_FUSED_THREADS / 32;
  const int ntile = (n + VS - 1) / VS;
  for (int i = tid; i < n; i += TRSV_FUSED_THREADS) w[i] = x[i];
  __syncthreads();
  for (int t = 0; t < ntile; ++t) {  // L z = b
    const int j0 = t * VS, nt = min(VS, n - j0);
    stage_inverse_block_wide(dinv_base + (long long)(t / 2) * TS * TS + (t % 2) * ((long long)VS * TS + VS), S, tid);
    if (tid < VS) yj[tid] = tid < nt ? w[j0 + tid] : 0.0;
    __syncthreads();
    if (tid < 4 * VS) {
      const int r = tid >> 2, part = tid & 3;
      double acc = 0.0;
#pragma unroll 4
      for (int k = part; k <= r; k += 4) acc = fma(S[r][k], yj[k], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) {
        zj[r] = acc;
        if (r < nt) z[j0 + r] = acc;
      }
    }
    __syncthreads();
    const double z0 = zj[lane], z1 = zj[lane + 32];
    for (int i0 = j0 + VS + warp * 4; i0 < n; i0 += NW * 4) {  // four rows per warp in flight
      double a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = min(i0 + u, n - 1);
        const double* row = L + (long long)i * ld + j0;
        a[u] = row[lane] * z0 + row[lane + 32] * z1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double acc = warp_sum(a[u]);
        if (lane == 0 && i0 + u < n) w[i0 + u] -= acc;
      }
    }
    __syncthreads();
  }
  for (int t = ntile - 1; t >= 0; --t) {  // L^T x = z
    const int j0 = t * VS, nt = min(VS, n - j0);
    stage_inverse_block_wide(dinv_base + (long long)(t / 2) * TS * TS + (t % 2) * ((long long)VS * TS + VS), S, tid);
    if (tid < VS) zj[tid] = tid < nt ? z[j0 + tid] : 0.0;
    __syncthreads();
    if (tid < 4 * VS) {
      const int c = tid >> 2, part = tid & 3;
      double acc = 0.0;
#pragma unroll 4
      for (int r = c + part; r < VS; r += 4) acc = fma(S[r][c], zj[r], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) {
        yj[c] = acc;
        if (c < nt) x[j0 + c] = acc;
      }
    }
    __syncthreads();
    for (int c = tid; c < j0; c += TRSV_FUSED_THREADS) {
      const double* col = L + (long long)j0 * ld + c;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int r = 0;
      for (; r + 8 <= nt; r += 8) {  // eight independent loads in flight; accumulation order of bwd_step_kernel
        const double v0 = col[(long long)r * ld], v1 = col[(long long)(r + 1) * ld], v2 = col[(long long)(r + 2) * ld],
                     v3 = col[(long long)(r + 3) * ld], v4 = col[(long long)(r + 4) * ld], v5 = col[(long long)(r + 5) * ld],
                     v6 = col[(long long)(r + 6) * ld], v7 = col[(long long)(r + 7) * ld];
        a0 = fma(v0, yj[r], a0), a1 = fma(v1, yj[r + 1], a1), a2 = fma(v2, yj[r + 2], a2), a3 = fma(v3, yj[r + 3], a3);
        a0 = fma(v4, yj[r + 4], a0), a1 = fma(v5, yj[r + 5], a1), a2 = fma(v6, yj[r + 6], a2), a3 = fma(v7, yj[r + 7], a3);
      }
      for (; r + 4 <= nt; r += 4) {
        a0 = fma(col[(long long)r * ld], yj[r], a0);
        a1 = fma(col[(long long)(r + 1) * ld], yj[r + 1], a1);
        a2 = fma(col[(long long)(r + 2) * ld], yj[r + 2], a2);
        a3 = fma(col[(long long)(r + 3) * ld], yj[r + 3], a3);
      }
      for (; r < nt; ++r) a0 = fma(col[(long long)r * ld], yj[r], a0);
      z[c] -= (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
  }
}

// Block-range variants of the fused substitution for n > TRSV_FUSED_MAX: ONE launch walks all 64-wide tiles of a
// VBLK-column diagonal block [b0, b1) (the step-per-launch chain was 782 launches per direction at N = 50 000, 32 ms for a
// 20 GB read; 4.4 ms at N = 8192).  Rows / columns outside the block are left to the full-grid panel GEMVs of potrs_few.
// Arithmetic per row / column is that of the step kernels (same accumulation order).
__global__ void __launch_bounds__(TRSV_FUSED_THREADS) trsv_block_fwd_kernel(const double* __restrict__ L, long long ld,
                                                                            int b0, int b1,
                                                                            const double* __restrict__ dinv_base, double* w,
                                                                            double* z, long long bstride) {
  L += blockIdx.x * bstride, dinv_base += blockIdx.x * bstride, w += blockIdx.x * bstride, z += blockIdx.x * bstride;
  __shared__ double S[VS][VSP];
  __shared__ double yj[VS], zj[VS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = TRSV_FUSED_THREADS / 32;
  for (int j0 = b0; j0 < b1; j0 += VS) {  // L z = b inside the block
    const int t = j0 / VS, nt = min(VS, b1 - j0);
    stage_inverse_block_wide(dinv_base + (long long)(t / 2) * TS * TS + (t % 2) * ((long long)VS * TS + VS), S, tid);
    if (tid < VS) yj[tid] = tid < nt ? w[j0 + tid] : 0.0;
    __syncthreads();
    if (tid < 4 * VS) {
      const int r = tid >> 2, part = tid & 3;
      double acc = 0.0;
#pragma unroll 4
      for (int k = part; k <= r; k += 4) acc = fma(S[r][k], yj[k], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) {
        zj[r] = acc;
        if (r < nt) z[j0 + r] = acc;
      }
    }
    __syncthreads();
    const double z0 = zj[lane], z1 = zj[lane + 32];
    for (int i0 = j0 + VS + warp * 4; i0 < b1; i0 += NW * 4) {  // four rows per warp in flight
      double a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = min(i0 + u, b1 - 1);
        const double* row = L + (long long)i * ld + j0;
        a[u] = row[lane] * z0 + row[lane + 32] * z1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double acc = warp_sum(a[u]);
        if (lane == 0 && i0 + u < b1) w[i0 + u] -= acc;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(TRSV_FUSED_THREADS) trsv_block_bwd_kernel(const double* __restrict__ L, long long ld,
                                                                            int b0, int b1,
                                                                            const double* __restrict__ dinv_base, double* z,
                                                                            double* x, long long bstride) {
  L += blockIdx.x * bstride, dinv_base += blockIdx.x * bstride, z += blockIdx.x * bstride, x += blockIdx.x * bstride;
  __shared__ double S[VS][VSP];
  __shared__ double yj[VS], zj[VS];
  const int tid = threadIdx.x;
  const int last = b0 + ((b1 - b0 - 1) / VS) * VS;
  for (int j0 = last; j0 >= b0; j0 -= VS) {  // L^T x = z inside the block
    const int t = j0 / VS, nt = min(VS, b1 - j0);
    stage_inverse_block_wide(dinv_base + (long long)(t / 2) * TS * TS + (t % 2) * ((long long)VS * TS + VS), S, tid);
    if (tid < VS) zj[tid] = tid < nt ? z[j0 + tid] : 0.0;
    __syncthreads();
    if (tid < 4 * VS) {
      const int c = tid >> 2, part = tid & 3;
      double acc = 0.0;
#pragma unroll 4
      for (int r = c + part; r < VS; r += 4) acc = fma(S[r][c], zj[r], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) {
        yj[c] = acc;
        if (c < nt) x[j0 + c] = acc;
      }
    }
    __syncthreads();
    for (int c = b0 + tid; c < j0; c += TRSV_FUSED_THREADS) {
      const double* col = L + (long long)j0 * ld + c;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int r = 0;
      for (; r + 8 <= nt; r += 8) {
        const double v0 = col[(long long)r * ld], v1 = col[(long long)(r + 1) * ld], v2 = col[(long long)(r + 2) * ld],
                     v3 = col[(long long)(r + 3) * ld], v4 = col[(long long)(r + 4) * ld], v5 = col[(long long)(r + 5) * ld],
                     v6 = col[(long long)(r + 6) * ld], v7 = col[(long long)(r + 7) * ld];
        a0 = fma(v0, yj[r], a0), a1 = fma(v1, yj[r + 1], a1), a2 = fma(v2, yj[r + 2], a2), a3 = fma(v3, yj[r + 3], a3);
        a0 = fma(v4, yj[r + 4], a0), a1 = fma(v5, yj[r + 5], a1), a2 = fma(v6, yj[r + 6], a2), a3 = fma(v7, yj[r + 7], a3);
      }
      for (; r + 4 <= nt; r += 4) {
        a0 = fma(col[(long long)r * ld], yj[r], a0);
        a1 = fma(col[(long long)(r + 1) * ld], yj[r + 1], a1);
        a2 = fma(col[(long long)(r + 2) * ld], yj[r + 2], a2);
        a3 = fma(col[(long long)(r + 3) * ld], yj[r + 3], a3);
      }
      for (; r < nt; ++r) a0 = fma(col[(long long)r * ld], yj[r], a0);
      z[c] -= (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) logdet_kernel(const double* __restrict__ L, long long ld, int n, double* out,
                                                      long long bstride) {
  L += blockIdx.x * bstride, out += blockIdx.x * bstride;
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) s += log(fabs(L[(long long)i * ld + i]));
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = 2.0 * s;
}

__global__ void __launch_bounds__(1024) dot_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                   long long n, double* out) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) s += a[i] * b[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}


// ----------------------------------------------------------------------------------------------
// Matrix-vector products for the block-cyclic (multi-GPU) triangular solves: the panels of L that a
// rank owns are m x n (n <= one block), row-major.  HBM-read bound (8*m*n bytes), deterministic.
// ----------------------------------------------------------------------------------------------
// y[i] += alpha * sum_j A[i][j] x[j]; one warp per row.
__global__ void __launch_bounds__(256) gemv_n_kernel(const double* __restrict__ A, long long lda, int m, int n,
                                                     double alpha, const double* __restrict__ x, double* y,
                                                     long long bstride) {
  A += blockIdx.y * bstride, x += blockIdx.y * bstride, y += blockIdx.y * bstride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long i = (long long)blockIdx.x * 8 + warp; i < m; i += (long long)gridDim.x * 8) {
    const double* row = A + i * lda;
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) acc = fma(row[j], x[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[i] += alpha * acc;
  }
}

// partial[chunk][j] = sum_{i in chunk} A[i][j] x[i]; thread = column, CTA = (column group, row chunk).
constexpr int GEMVT_ROWS = 256;
__global__ void __launch_bounds__(256) gemv_t_partial_kernel(const double* __restrict__ A, long long lda, int m, int n,
                                                             const double* __restrict__ x, double* partial,
                                                             long long bstride) {
  A += blockIdx.z * bstride, x += blockIdx.z * bstride, partial += blockIdx.z * bstride;
  __shared__ double xs[GEMVT_ROWS];
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int i0 = blockIdx.y * GEMVT_ROWS, i1 = min(m, i0 + GEMVT_ROWS);
  if (i0 + (int)threadIdx.x < i1) xs[threadIdx.x] = x[i0 + threadIdx.x];
  __syncthreads();
  if (j >= n) return;
  double a0 = 0.0, a1 = 0.0;
  int i = i0;
  for (; i + 1 < i1; i += 2) {
    a0 = fma(A[(long long)i * lda + j], xs[i - i0], a0);
    a1 = fma(A[(long long)(i + 1) * lda + j], xs[i + 1 - i0], a1);
  }
  if (i < i1) a0 = fma(A[(long long)i * lda + j], xs[i - i0], a0);
  partial[(long long)blockIdx.y * n + j] = a0 + a1;
}

__global__ void __launch_bounds__(256) gemv_t_reduce_kernel(const double* __restrict__ partial, int chunks, int n,
                                                            double alpha, double* y, long long bstride) {
  partial += blockIdx.y * bstride, y += blockIdx.y * bstride;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  double s = 0.0;
  for (int c = 0; c < chunks; ++c) s += partial[(long long)c * n + j];
  y[j] += alpha * s;
}

// ----------------------------------------------------------------------------------------------
// Host-side recursion
// ----------------------------------------------------------------------------------------------
static int g_ozaki_all = -1;               // 1: INT8-slice POTRF updates also with block columns narrower than 2048 (N < 40 000)
static int g_ozaki_potri_min_n = 40000;   // smallest N whose POTRI uses the INT8-slice path at all (fvgp_set_ozaki_gate)
static int g_ozaki_lauum_min = 8192;      // smallest P11 (rows) whose SYRK update inside LAUUM goes through the INT8 path
constexpr int OZAKI_MIN_CHUNK = 1536;      // fewest rows per chunk of a triangular INT8 product (multiple of BM)
constexpr int64_t OZAKI_NBLOCK = 4096;     // column block of the INT8-slice GEMM's int32 planes

struct Ctx {
  cudaStream_t st;
  double* dinv;        // per-64-tile inverses, tile t at dinv + t*4096
  int* info;
  double* work;        // scratch panel for trtri / lauum
  int err;
  int batch = 1;       // > 1: every launch covers `batch` problems of identical shape (population evaluation) whose
  long long bstride = 0;  // buffers (matrix, tile inverses, scratch) sit `bstride` doubles apart; info is an int per problem
  // INT8-slice SYRK inside LAUUM (fvgp_potri_lower): scratch for the transposed panel + the slices, 0 slices = off
  int oz_slices = 0;
  void* oz_work = nullptr;
  long long oz_bytes = 0;
  int oz_tri = 0;      // > 0: also the products with a triangular operand, contraction range cut into this many chunks
};

static inline int split(int n) {  // n > TS: first part is a multiple of 128, at least 128, less than n
  int h = ((n / 2 + BM / 2) / BM) * BM;
  if (h < BM) h = BM;
  if (h >= n) h -= BM;
  return h;
}

#define REC_OK(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// B (m x n) <- B * L^-T, L the n x n lower factor whose first row is global row `row0`.
static int trsm_rt_rec(Ctx& c, double* B, long long ldb, int m, const double* L, long long ld, int n, int row0) {
  if (m <= 0) return 0;
  if (n <= TS) {
    const double* tile = c.dinv + (long long)(row0 / TS) * TS * TS;
    return launch_gemm<false, false>(c.st, B, ldb, tile, TS, B, ldb, m, n, n, 1.0, 0.0, 0, c.batch, c.bstride);
  }
  const int n1 = split(n), n2 = n - n1;
  REC_OK(trsm_rt_rec(c, B, ldb, m, L, ld, n1, row0));
  REC_OK((launch_gemm<false, false>(c.st, B, ldb, L + (long long)n1 * ld, ld, B + n1, ldb, m, n2, n1, -1.0, 1.0, 0, c.batch,
                                    c.bstride)));
  return trsm_rt_rec(c, B + n1, ldb, m, L + (long long)n1 * ld + n1, ld, n2, row0 + n1);
}

// B (m x n) <- B * L^-1 (the L^T x = z direction for row-stored right-hand sides).
static int trsm_rn_rec(Ctx& c, double* B, long long ldb, int m, const double* L, long long ld, int n, int row0) {
  if (m <= 0) return 0;
  if (n <= TS) {
    const double* tile = c.dinv + (long long)(row0 / TS) * TS * TS;
    return launch_gemm<false, true>(c.st, B, ldb, tile, TS, B, ldb, m, n, n, 1.0, 0.0, 0, c.batch, c.bstride);
  }
  const int n1 = split(n), n2 = n - n1;
  REC_OK(trsm_rn_rec(c, B + n1, ldb, m, L + (long long)n1 * ld + n1, ld, n2, row0 + n1));
  REC_OK((launch_gemm<false, true>(c.st, B + n1, ldb, L + (long long)n1 * ld, ld, B, ldb, m, n1, n2, -1.0, 1.0, 0, c.batch,
                                   c.bstride)));
  return trsm_rn_rec(c, B, ldb, m, L, ld, n1, row0);
}

static int potrf_rec(Ctx& c, double* A, long long ld, int n, int row0) {
  if (n <= TS) {
    double* tile_inv = c.dinv + (long long)(row0 / TS) * TS * TS;
    if (use_tile_v1())
      launch(potrf_tile_kernel, c.batch, 512, POTRF_TILE_SMEM, c.st, A, ld, n, tile_inv, c.info, row0, c.bstride);
    else if (use_tile_v2())
      launch(potrf_tile2_kernel, c.batch, 512, POTRF_TILE2_SMEM, c.st, A, ld, n, tile_inv, c.info, row0, c.bstride);
    else
      launch(potrf_tile3_kernel, c.batch, 512, POTRF_TILE3_SMEM, c.st, A, ld, n, tile_inv, c.info, row0, c.bstride);
    FVGP_LAUNCH_OK();
    return 0;
  }
  const int n1 = split(n), n2 = n - n1;
  double* A21 = A + (long long)n1 * ld;
  double* A22 = A21 + n1;
  REC_OK(potrf_rec(c, A, ld, n1, row0));
  REC_OK(trsm_rt_rec(c, A21, ld, n2, A, ld, n1, row0));
  REC_OK((launch_gemm<false, false>(c.st, A21, ld, A21, ld, A22, ld, n2, n2, n1, -1.0, 1.0, GEMM_LOWER, c.batch, c.bstride)));
  return potrf_rec(c, A22, ld, n2, row0 + n1);
}

// INT8-slice (Ozaki) trailing updates: 0 = off (DMMA), S >= 6 = number of 6-bit slices (csrc/ozaki.cu).  Set by
// fvgp_set_ozaki or the FVGP_OZAKI environment variable; used for trailing updates with at least OZAKI_MIN_M rows.
static int g_ozaki_slices = -1;
constexpr int OZAKI_MIN_M = 8192;
static int ozaki_slices() {
  if (g_ozaki_slices < 0) {
    // default ON with 8 slices (FVGP_OZAKI=0 switches it off): LML / gradient within 1.6e-12 / 2.7e-11 of the DMMA path
    // and inside the 1e-8 oracle parity at N = 8000 / 16 000 / 50 000; POTRF at N = 50 000 1.305 -> 0.96 s
    const char* e = getenv("FVGP_OZAKI");
    int v = e ? atoi(e) : 8;
    if (v == 1) v = 8;
    g_ozaki_slices = (v >= 6 && v <= 10 && fvgp_ozaki_available()) ? v : 0;
  }
  return g_ozaki_slices;
}

// Chunk count of the INT8-slice products with a triangular operand inside POTRI (0 = those stay on DMMA).  Default 8:
// POTRI at N = 50 000 2.30 s (SYRK half only) -> 1.97 (4 chunks) / 1.67 (8) / 1.73 (12) / 1.76 (16) / 1.87 s (24),
// gradient within 7e-10 of the all-DMMA path at every setting (profiles/r02/ozaki_tri_probe.v15.log, .v16.log).
static int g_ozaki_tri = -1;
static int ozaki_tri_chunks() {
  if (g_ozaki_tri < 0) {
    const char* e = getenv("FVGP_OZAKI_TRI");
    const int v = e ? atoi(e) : 8;
    g_ozaki_tri = v < 0 ? 0 : (v > 32 ? 32 : v);
  }
  return g_ozaki_tri;
}

// The INT8 scratch comes from the stream-ordered pool, which by default returns freed memory to the OS at the next
// synchronisation: the 7.5 GB of an N = 50 000 factorisation then came back through the driver on every call (0.3 ... 1.7 s
// each, the "lottery" of profiles/r02/ozaki_step_probe.v10.log).  While the matrices leave room the pool keeps what it
// has; when they do not (two N x N buffers + workspace beyond ~40 % of the device: N ~ 100 000 on one GPU) cached scratch
// would starve the caller's own allocator (torch ran out of memory for the POTRI workspace with 19 GB idle in the pool:
// profiles/r02/v26_c3_single_gpu_probe.log), so the pool goes back to releasing and is trimmed at once.
static void pool_retention(long long n) {
  int dev = 0;
  cudaMemPool_t pool;
  size_t free_b = 0, total_b = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess ||
      cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  const int roomy = 2.5 * 8.0 * (double)n * (double)n < (double)total_b ? 1 : 0;
  static int state = -1;
  if (state != roomy) {
    unsigned long long keep = roomy ? ~0ull : 0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    if (!roomy) cudaMemPoolTrimTo(pool, 0);
    cudaGetLastError();
    state = roomy;
  }
}

// Right-looking blocked Cholesky with one step of look-ahead on two streams.
//
// The recursion above is flop-optimal but strictly serial: the latency-bound work on the diagonal (tile
// factorisations, 1-CTA triangular products; ~33 ms per 8192 rows) sits between the large trailing updates
// and leaves the tensor pipe idle ~15 % of the time at N = 50 000.  Here block column k+1 is updated,
// factored and solved on a high-priority side stream (P) while the caller's stream (S) is still busy with the
// trailing update of step k:
//     P: panel(0)
//     step k:  S: wait panel(k);  U(k):  A[k+2:, k+2:] -= L[k+2:, k] L[k+2:, k]^T      (SYRK, lower tiles)
//              P: wait U(k-1);    LA(k): A[k+1:, k+1]  -= L[k+1:, k] L[k+1, k]^T       (one block column)
//                                 panel(k+1): potrf(A[k+1, k+1]) (recursive), A[k+2:, k+1] <- A[k+2:, k+1] L^-T
// U(k) and LA(k)/panel(k+1) touch disjoint blocks.  Streams and events are created per call (re-entrant).
static int potrf_lookahead(Ctx& c, double* A, long long ld, int n, int nb) {
  const int nblk = (n + nb - 1) / nb;
  cudaStream_t S = c.st, P;
  int lo_prio = 0, hi_prio = 0;
  FVGP_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
  FVGP_CUDA_OK(cudaStreamCreateWithPriority(&P, cudaStreamNonBlocking, hi_prio));
  std::vector<cudaEvent_t> ev_panel(nblk), ev_trail(nblk);
  for (int k = 0; k < nblk; ++k) {
    FVGP_CUDA_OK(cudaEventCreateWithFlags(&ev_panel[k], cudaEventDisableTiming));
    FVGP_CUDA_OK(cudaEventCreateWithFlags(&ev_trail[k], cudaEventDisableTiming));
  }
  auto blk = [&](int i, int j) { return A + (long long)i * nb * ld + (long long)j * nb; };
  auto rows_from = [&](int i) { return n - i * nb; };
  Ctx cp = c;
  cp.st = P;
  int rc = 0;
  // INT8-slice trailing updates: one scratch allocation for the largest update of this factorisation
  // (measured with 2048-wide block columns only, i.e. N >= 40 000; smaller N stay on the DMMA pipe unless FVGP_OZAKI_ALL=1)
  if (g_ozaki_all < 0) {
    const char* e = getenv("FVGP_OZAKI_ALL");
    g_ozaki_all = (e && atoi(e) == 1) ? 1 : 0;
  }
  const int oz_all = g_ozaki_all;
  const int oz = (nb % 16 == 0 && (nb >= 2048 || oz_all) && n - 2 * nb >= OZAKI_MIN_M) ? ozaki_slices() : 0;
  void *oz_work = nullptr, *ozp_work = nullptr;  // scratch of the updates on S / of the look-ahead column on P
  int64_t oz_bytes = 0, ozp_bytes = 0;
  if (oz) {
    oz_bytes = fvgp_ozaki_work_bytes(n - 2 * nb, n - 2 * nb, nb, oz, OZAKI_NBLOCK);
    pool_retention(n);
    const cudaError_t me = cudaMallocAsync(&oz_work, (size_t)oz_bytes, S);
    if (me != cudaSuccess) {
      cudaGetLastError();
      oz_work = nullptr;  // not enough memory: stay on the DMMA path
    }
    // FVGP_OZAKI_PANEL=0: keep the look-ahead column update (m x nb x nb, panel stream) on the DMMA pipe
    static int oz_panel = -1;
    if (oz_panel < 0) {
      const char* e = getenv("FVGP_OZAKI_PANEL");
      oz_panel = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (oz_work != nullptr && oz_panel) {
      ozp_bytes = fvgp_ozaki_work_bytes(n - 2 * nb, nb, nb, oz, OZAKI_NBLOCK);
      if (cudaMallocAsync(&ozp_work, (size_t)ozp_bytes, S) != cudaSuccess) {
        cudaGetLastError();
        ozp_work = nullptr;
      }
    }
    static bool told = false;
    if (!told) {
      told = true;
      fprintf(stderr, "[fvgp_b200] potrf n=%d nb=%d: INT8-slice trailing updates, %d slices, scratch %.2f + %.2f GB: %s\n", n, nb,
              oz, oz_bytes / 1e9, ozp_work ? ozp_bytes / 1e9 : 0.0, oz_work ? "on" : cudaGetErrorString(me));
    }
  }
  // FVGP_OZAKI_STREAMS=1: issue everything on one stream.  Default 2 (panel path on its own high-priority stream next to
  // the INT8 updates): N = 50 000 POTRF 836 vs 898 ms with the CTA-pair tile (profiles/r02/potrf_50k_ozaki_variants.v12.log).
  static int oz_streams = -1;
  if (oz_streams < 0) {
    const char* e = getenv("FVGP_OZAKI_STREAMS");
    oz_streams = (e && atoi(e) == 1) ? 1 : 2;
  }
  if (oz_work != nullptr && oz_streams == 1) {
    // one-stream order (A/B): update, look-ahead column, panel -- no overlap between the persistent int8 GEMMs and the
    // latency-bound panel path
    cudaStreamDestroy(P);
    P = S;
    cp.st = S;
  }
  auto panel = [&](int k) -> int {  // factor the diagonal block k, solve the blocks below it
    const int w = std::min(nb, rows_from(k));
    REC_OK(potrf_rec(cp, blk(k, k), ld, w, k * nb));
    if (k + 1 < nblk) REC_OK(trsm_rt_rec(cp, blk(k + 1, k), ld, rows_from(k + 1), blk(k, k), ld, w, k * nb));
    FVGP_CUDA_OK(cudaEventRecord(ev_panel[k], P));
    return 0;
  };
  // the side stream starts after everything already queued on the caller's stream (the K-fill)
  FVGP_CUDA_OK(cudaEventRecord(ev_trail[nblk - 1], S));
  FVGP_CUDA_OK(cudaStreamWaitEvent(P, ev_trail[nblk - 1], 0));
  rc = panel(0);
  for (int k = 0; rc == 0 && k + 1 < nblk; ++k) {
    // S: trailing update of the blocks beyond column k+1
    FVGP_CUDA_OK(cudaStreamWaitEvent(S, ev_panel[k], 0));
    if (k + 2 < nblk) {
      const int m2 = rows_from(k + 2);
      bool done = false;
      if (oz_work != nullptr && m2 >= OZAKI_MIN_M) {
        // the INT8 path validates its arguments before it touches C: a refusal falls back to the DMMA update for good
        const int orc = fvgp_ozaki_gemm_nt(blk(k + 2, k + 2), ld, blk(k + 2, k), ld, blk(k + 2, k), ld, m2, m2, nb, -1.0, 1, 0,
                                           1, oz, OZAKI_NBLOCK, oz_work, oz_bytes, S);
        done = orc == 0;
        if (orc == -100) {  // failed after part of the block had been updated: not recoverable
          rc = FVGP_ERR_CUDA;
          break;
        }
        if (!done) {
          fprintf(stderr, "[fvgp_b200] potrf: INT8-slice update refused (m = %d); continuing on the DMMA pipe\n", m2);
          g_ozaki_slices = 0;
        }
      }
      if (!done)
        rc = launch_gemm<false, false>(S, blk(k + 2, k), ld, blk(k + 2, k), ld, blk(k + 2, k + 2), ld, m2, m2, nb, -1.0, 1.0,
                                       GEMM_LOWER);
      if (rc != 0) break;
    }
    FVGP_CUDA_OK(cudaEventRecord(ev_trail[k], S));
    // P: look-ahead update of block column k+1 (needs U(k-1), which touched that column), then its panel
    if (k > 0) FVGP_CUDA_OK(cudaStreamWaitEvent(P, ev_trail[k - 1], 0));
    const int w1 = std::min(nb, rows_from(k + 1));
    rc = launch_gemm<false, false>(P, blk(k + 1, k), ld, blk(k + 1, k), ld, blk(k + 1, k + 1), ld, w1, w1, nb, -1.0, 1.0,
                                   GEMM_LOWER);
    if (rc == 0 && k + 2 < nblk) {
      const int m2 = rows_from(k + 2);
      bool done = false;
      if (ozp_work != nullptr && g_ozaki_slices > 0 && m2 >= OZAKI_MIN_M && w1 == nb) {
        // the block column below the diagonal block (m2 x nb x nb): the same INT8 route, own scratch on the panel stream
        const int orc = fvgp_ozaki_gemm_nt(blk(k + 2, k + 1), ld, blk(k + 2, k), ld, blk(k + 1, k), ld, m2, w1, nb, -1.0, 0, 0, 0,
                                           oz, OZAKI_NBLOCK, ozp_work, ozp_bytes, P);
        if (orc == -100) {
          rc = FVGP_ERR_CUDA;
          break;
        }
        done = orc == 0;
      }
      if (!done)
        rc = launch_gemm<false, false>(P, blk(k + 2, k), ld, blk(k + 1, k), ld, blk(k + 2, k + 1), ld, m2, w1, nb, -1.0, 1.0, 0);
    }
    if (rc == 0) rc = panel(k + 1);
  }
  // the caller's stream continues only after the last panel
  if (rc == 0) {
    cudaEventRecord(ev_panel[nblk - 1], P);
    cudaStreamWaitEvent(S, ev_panel[nblk - 1], 0);
  } else {
    cudaStreamSynchronize(P);
  }
  for (int k = 0; k < nblk; ++k) {
    cudaEventDestroy(ev_panel[k]);
    cudaEventDestroy(ev_trail[k]);
  }
  if (oz_work != nullptr) cudaFreeAsync(oz_work, S);
  if (ozp_work != nullptr) cudaFreeAsync(ozp_work, S);  // after ev_panel[last]: S waits for everything queued on P
  if (P != S) cudaStreamDestroy(P);
  return rc;
}

// Block width of the look-ahead factorisation; 0 disables it (pure recursion).  FVGP_POTRF_NB overrides.
// Measured on B200 (profiles/r01/potrf_sweep.v4.log): N = 50 000: recursion 1468 ms, nb 1024 1472 ms, nb 2048
// 1330 ms, nb 4096 1357 ms; N = 16 384: recursion 103 ms, nb 1024 89 ms, nb 2048 114 ms.
static int potrf_block_width(int n) {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("FVGP_POTRF_NB");
    env = e ? atoi(e) : -1;
    if (env > 0) env = std::max(BM, (env / BM) * BM);
  }
  if (env >= 0) return (env > 0 && n >= 4 * env) ? env : 0;
  // round 2 sweep (profiles/r02/potrf_nb_sweep.v11.log, ms): N = 8192: recursion 18.2, nb 512 14.4, 1024 16.3, 2048 17.6;
  // N = 12 288: 42.4 / 31.6 / 42.3 / 37.0;  N = 16 384: 73.2 / 62.6 / 61.0 / 72.3;  N = 24 576: 197.8 / 202.6 / 177.1 / 189.8
  if (n >= 40000) return 2048;
  if (n >= 14336) return 1024;
  if (n >= 6144) return 512;
  return 0;
}

// L -> L^-1 in place (lower).  Needs explicit zeros above the diagonal inside diagonal blocks.
// ---- INT8-slice versions of the POTRI products that have a TRIANGULAR operand (csrc/ozaki.cu; opt-in, Ctx::oz_tri).
// A full-K int8 GEMM would do twice the useful work, so the output is cut into `chunks` blocks along the triangular
// operand and every block only multiplies the part of the contraction range that operand reaches: (c + 1) / c times
// the useful work at c chunks (25 % extra at 4, 12.5 % at 8) instead of 2x.  All products are brought into the NT form of fvgp_ozaki_gemm_nt by
// transposing the triangular operand into scratch with its never-written upper triangle masked to zero.
// Return 0 = done, 1 = refused before anything was modified (caller falls back to the DMMA products), < 0 = error.
static inline long long align256(long long b) { return (b + 255) / 256 * 256; }

static int oz_chunk(int n, int chunks) {
  int cb = ((n + chunks - 1) / chunks + BM - 1) / BM * BM;
  return cb < OZAKI_MIN_CHUNK ? OZAKI_MIN_CHUNK : cb;  // short row blocks leave the 256 x 256 tiles' last wave half empty
}

// TRTRI: L21 <- -M22 L21 M11 (M11 = L[0:n1,0:n1], M22 already inverted in place, both lower).
static int trtri_products_int8(Ctx& c, double* L, long long ld, int n1, int n2) {
  double* L21 = L + (long long)n1 * ld;
  double* L22 = L21 + n1;
  const int cb1 = oz_chunk(n1, c.oz_tri), cb2 = oz_chunk(n2, c.oz_tri);
  const long long bt_bytes = align256((long long)n1 * n1 * 8), ab_bytes = align256((long long)cb2 * n2 * 8);
  const long long head = bt_bytes > ab_bytes ? bt_bytes : ab_bytes;
  const long long need = head + fvgp_ozaki_work_bytes(cb1 > cb2 ? cb1 : cb2, n1 > n2 ? n1 : n2, n1 > n2 ? n1 : n2, c.oz_slices,
                                                       OZAKI_NBLOCK);
  if (need > c.oz_bytes || n2 % 16 != 0) return 1;
  double* Bt = (double*)c.oz_work;  // Bt[j][k] = M11[k][j] for k >= j, 0 otherwise
  void* ow = (char*)c.oz_work + head;
  const long long ob = c.oz_bytes - head;
  double* Wt = c.work;  // n1 x n2: Wt = M11^T L21^T
  launch(transpose_pad_kernel, dim3((n1 + 31) / 32, (n1 + 31) / 32), 256, 0, c.st, Bt, (long long)n1, (const double*)L, ld, n1, n1,
         n1, 1);
  FVGP_LAUNCH_OK();
  FVGP_CUDA_OK(cudaMemsetAsync(Wt, 0, (size_t)n1 * n2 * sizeof(double), c.st));
  for (int jb = 0; jb < n1; jb += cb1) {  // rows [jb, je) of Wt only see k >= jb
    const int je = std::min(jb + cb1, n1);
    const int rc = fvgp_ozaki_gemm_nt(Wt + (long long)jb * n2, n2, Bt + (long long)jb * n1 + jb, n1, L21 + jb, ld, je - jb, n2,
                                      n1 - jb, 1.0, 0, 0, 0, c.oz_slices, OZAKI_NBLOCK, ow, ob, c.st);
    if (rc != 0) return (jb == 0 && rc != -100) ? 1 : FVGP_ERR_CUDA;  // only scratch was written so far
  }
  // L21 <- -M22 Wt^T: row block [ib, ie) only sees k < ie; the block of M22 goes through a zero-masked copy
  double* Ab = (double*)c.oz_work;
  FVGP_CUDA_OK(cudaMemset2DAsync(L21, ld * sizeof(double), 0, (size_t)n1 * sizeof(double), n2, c.st));
  for (int ib = 0; ib < n2; ib += cb2) {
    const int ie = std::min(ib + cb2, n2);
    launch(copy_lower_rows_kernel, dim3((ie + 255) / 256, (ie - ib + 15) / 16), 256, 0, c.st, Ab, (long long)ie,
           (const double*)(L22 + (long long)ib * ld), ld, ie - ib, ie, ib);
    FVGP_LAUNCH_OK();
    const int rc = fvgp_ozaki_gemm_nt(L21 + (long long)ib * ld, ld, Ab, ie, Wt, n2, ie - ib, n1, ie, -1.0, 0, 0, 0, c.oz_slices,
                                      OZAKI_NBLOCK, ow, ob, c.st);
    if (rc != 0) return FVGP_ERR_CUDA;  // L21 has been overwritten: no way back
  }
  return 0;
}

// LAUUM: W = M22^T M21 (n2 x n1) into c.work, given T = M21^T (n1 x k16, zero-padded) already in scratch at `T`.
static int lauum_w_int8(Ctx& c, const double* M22, long long ld, int n1, int n2, const double* T, long long k16, void* free_work,
                        long long free_bytes) {
  const int cb = oz_chunk(n2, c.oz_tri);
  const long long u_bytes = align256((long long)n2 * n2 * 8);
  if (u_bytes + fvgp_ozaki_work_bytes(cb, n1, n2, c.oz_slices, OZAKI_NBLOCK) > free_bytes || n2 % 16 != 0) return 1;
  double* U = (double*)free_work;  // U[i][k] = M22[k][i] for k >= i, 0 otherwise
  void* ow = (char*)free_work + u_bytes;
  const long long ob = free_bytes - u_bytes;
  launch(transpose_pad_kernel, dim3((n2 + 31) / 32, (n2 + 31) / 32), 256, 0, c.st, U, (long long)n2, M22, ld, n2, n2, n2, 1);
  FVGP_LAUNCH_OK();
  FVGP_CUDA_OK(cudaMemsetAsync(c.work, 0, (size_t)n2 * n1 * sizeof(double), c.st));
  for (int ib = 0; ib < n2; ib += cb) {  // rows [ib, ie) of W only see k >= ib
    const int ie = std::min(ib + cb, n2);
    const int rc = fvgp_ozaki_gemm_nt(c.work + (long long)ib * n1, n1, U + (long long)ib * n2 + ib, n2, T + ib, k16, ie - ib, n1,
                                      n2 - ib, 1.0, 0, 0, 0, c.oz_slices, OZAKI_NBLOCK, ow, ob, c.st);
    if (rc != 0) return (ib == 0 && rc != -100) ? 1 : FVGP_ERR_CUDA;
  }
  return 0;
}

static int trtri_rec(Ctx& c, double* L, long long ld, int n, int row0) {
  if (n <= TS) {
    const double* tile = c.dinv + (long long)(row0 / TS) * TS * TS;
    launch(copy2d_kernel, dim3(1, (n + 15) / 16, c.batch), 128, 0, c.st, L, ld, tile, TS, n, n, c.bstride);
    FVGP_LAUNCH_OK();
    return 0;
  }
  const int n1 = split(n), n2 = n - n1;
  double* L21 = L + (long long)n1 * ld;
  double* L22 = L21 + n1;
  REC_OK(trtri_rec(c, L, ld, n1, row0));
  REC_OK(trtri_rec(c, L22, ld, n2, row0 + n1));
  if (c.oz_tri > 0 && c.oz_slices > 0 && c.batch == 1 && n1 >= g_ozaki_lauum_min && n2 >= g_ozaki_lauum_min / 2) {
    const int rc = trtri_products_int8(c, L, ld, n1, n2);
    if (rc <= 0) return rc;  // done, or failed for good; 1 = refused untouched: DMMA products below
  }
  // W = L21 * M11   (M11 lower: k >= column)
  REC_OK((launch_gemm<false, true>(c.st, L21, ld, L, ld, c.work, n1, n2, n1, n1, 1.0, 0.0, GEMM_KB_FROM_N, c.batch, c.bstride)));
  // L21 = -M22 * W  (M22 lower: k <= row)
  return launch_gemm<false, true>(c.st, L22, ld, c.work, n1, L21, ld, n2, n1, n2, -1.0, 0.0, GEMM_KE_FROM_M, c.batch, c.bstride);
}

// lower(M) <- lower(M^T M) in place.
static int lauum_rec(Ctx& c, double* M, long long ld, int n) {
  if (n <= BM) return launch_gemm<true, true>(c.st, M, ld, M, ld, M, ld, n, n, n, 1.0, 0.0, 0, c.batch, c.bstride);
  const int n1 = split(n), n2 = n - n1;
  double* M21 = M + (long long)n1 * ld;
  double* M22 = M21 + n1;
  REC_OK(lauum_rec(c, M, ld, n1));
  // P11 += M21^T M21
  bool syrk_done = false, w_done = false;
  if (c.oz_slices > 0 && c.batch == 1 && n1 >= g_ozaki_lauum_min && n2 >= 1024) {
    // INT8-slice path: the product is T T^T with T = M21^T (n1 x n2): transposed into scratch (k padded to 16 with zero
    // columns), then the same SYRK as the POTRF trailing update.  No triangular operand here, so nothing is wasted.
    const long long k16 = (n2 + 15) / 16 * 16;
    const long long t_bytes = ((long long)n1 * k16 * (long long)sizeof(double) + 255) / 256 * 256;
    // the contraction range is cut into `parts` pieces when the slices of the whole range do not fit the scratch
    // (N ~ 100 000 on one GPU: 40 GB of slices next to the 80 GB matrix); every piece accumulates into P11
    int parts = 1;
    long long kc = k16;
    while (parts < 32 && t_bytes + fvgp_ozaki_work_bytes(n1, n1, kc, c.oz_slices, OZAKI_NBLOCK) > c.oz_bytes) {
      parts *= 2;
      kc = ((k16 + parts - 1) / parts + 15) / 16 * 16;
    }
    if (t_bytes + fvgp_ozaki_work_bytes(n1, n1, kc, c.oz_slices, OZAKI_NBLOCK) <= c.oz_bytes) {
      double* T = (double*)c.oz_work;
      launch(transpose_pad_kernel, dim3((n1 + 31) / 32, (unsigned)((k16 + 31) / 32)), 256, 0, c.st, T, k16, (const double*)M21, ld,
             n2, n1, (int)k16, 0);
      FVGP_LAUNCH_OK();
      int orc = 0;
      for (long long k0 = 0; k0 < k16 && orc == 0; k0 += kc) {
        orc = fvgp_ozaki_gemm_nt(M, ld, T + k0, k16, T + k0, k16, n1, n1, std::min(kc, k16 - k0), 1.0, 1, 0, 1, c.oz_slices,
                                 OZAKI_NBLOCK, (char*)c.oz_work + t_bytes, c.oz_bytes - t_bytes, c.st);
        if (orc != 0 && k0 > 0) orc = -100;  // an earlier piece is already in P11
      }
      if (orc == -100) return FVGP_ERR_CUDA;
      syrk_done = orc == 0;
      if (!syrk_done) {
        c.oz_slices = 0;  // refused before C was touched: DMMA from here on
      } else if (c.oz_tri > 0 && n2 >= g_ozaki_lauum_min / 2) {
        const int wrc = lauum_w_int8(c, M22, ld, n1, n2, T, k16, (char*)c.oz_work + t_bytes, c.oz_bytes - t_bytes);
        if (wrc < 0) return wrc;
        w_done = wrc == 0;
      }
    }
  }
  if (!syrk_done)
    REC_OK((launch_gemm<true, true>(c.st, M21, ld, M21, ld, M, ld, n1, n1, n2, 1.0, 1.0, GEMM_LOWER, c.batch, c.bstride)));
  // W = M22^T M21  (M22 lower: k >= row of the output)
  if (!w_done)
    REC_OK((launch_gemm<true, true>(c.st, M22, ld, M21, ld, c.work, n1, n2, n1, n2, 1.0, 0.0, GEMM_KB_FROM_M, c.batch, c.bstride)));
  launch(copy2d_kernel, dim3((n1 + 255) / 256, (n2 + 15) / 16, c.batch), 256, 0, c.st, M21, ld, c.work, n1, n2, n1, c.bstride);
  FVGP_LAUNCH_OK();
  return lauum_rec(c, M22, ld, n2);
}

}  // namespace fvgp

using namespace fvgp;

// kfill.cu: gradient traces without host synchronisation (result in d_out on the device)
int trace_radial_enqueue(int kind, const double* d_x, int64_t n, int dim, const double* h_coord_scale,
                         const double* h_out_scale, const double* d_Kinv, int64_t ld, const double* d_b,
                         double* d_partials, double* d_out, cudaStream_t st);

// kfill.cu: lower-triangle fills of a batch of proposals in one launch (theta scratch on the device)
int64_t kfill_batch_theta_len(int batch);
int kfill_lower_batch_enqueue(int kind, const double* d_x, int64_t n, int dim, int batch, const double* h_amp,
                              const double* h_inv_scale, const double* h_length, const double* h_centre,
                              const double* d_noise, double* d_K, int64_t ldk, int64_t kstride, double* d_thetas,
                              cudaStream_t st);

static inline int64_t round_up16(int64_t v) { return (v + 15) / 16 * 16; }

extern "C" {

int fvgp_version(void) { return 100; }

// 0: DMMA trailing updates (default); 6..10: INT8-slice trailing updates with that many 6-bit slices (8 = FP64-grade
// for the LML); returns the previous setting.  Ignored when the library was built without CUTLASS.
int fvgp_set_ozaki(int slices) {
  const int old = ozaki_slices();
  g_ozaki_slices = (slices >= 6 && slices <= 10 && fvgp_ozaki_available()) ? slices : 0;
  return old;
}

int fvgp_set_ozaki_gate(int min_n, int min_rows) {
  if (min_n < 0 || min_rows < 2 * BM) return FVGP_ERR_ARG;
  g_ozaki_potri_min_n = min_n;
  g_ozaki_lauum_min = min_rows;
  g_ozaki_all = min_n < 40000 ? 1 : 0;
  return 0;
}

int fvgp_set_ozaki_tri(int chunks) {
  const int old = ozaki_tri_chunks();
  g_ozaki_tri = chunks < 0 ? 0 : (chunks > 32 ? 32 : chunks);
  return old;
}

unsigned long long fvgp_launch_count(void) { return g_launches; }

int64_t fvgp_chol_workspace_len(int64_t n) { return ((n + TS - 1) / TS) * (int64_t)TS * TS; }

int64_t fvgp_potrs_work_len(int64_t n) { return 2 * n + (VBLK / 256) * n + 64; }

int64_t fvgp_potri_workspace_len(int64_t n) {
  const int64_t h = n / 2 + BM + 2;
  return h * h;
}

static int potrf_lower_impl(double* d_A, int64_t n, int64_t lda, double* d_tileinv, int* d_info, void* stream, bool readback) {
  FVGP_REQUIRE(n > 0 && n < (1ll << 31) && lda >= n && lda % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c{st, d_tileinv, d_info, nullptr, 0};
  static bool configured = false;
  if (!configured) {
    FVGP_CUDA_OK(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)POTRF_TILE_SMEM));
    FVGP_CUDA_OK(cudaFuncSetAttribute(potrf_tile2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)POTRF_TILE2_SMEM));
    FVGP_CUDA_OK(cudaFuncSetAttribute(potrf_tile3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)POTRF_TILE3_SMEM));
    configured = true;
  }
  FVGP_CUDA_OK(cudaMemsetAsync(d_info, 0, sizeof(int), st));
  const int nb = potrf_block_width((int)n);
  int r = nb > 0 ? potrf_lookahead(c, d_A, lda, (int)n, nb) : potrf_rec(c, d_A, lda, (int)n, 0);
  if (r != 0 || !readback) return r;
  int info = 0;
  FVGP_CUDA_OK(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  FVGP_CUDA_OK(cudaStreamSynchronize(st));
  return info;
}

int fvgp_potrf_lower(double* d_A, int64_t n, int64_t lda, double* d_tileinv, int* d_info, void* stream) {
  return potrf_lower_impl(d_A, n, lda, d_tileinv, d_info, stream, true);
}

// Same factorisation WITHOUT the host synchronisation: the status (0 or the 1-based failing pivot) stays in *d_info for
// the caller to read when it next synchronises anyway (block-cyclic factorisation: one read for all panels).
int fvgp_potrf_lower_enqueue(double* d_A, int64_t n, int64_t lda, double* d_tileinv, int* d_info, void* stream) {
  return potrf_lower_impl(d_A, n, lda, d_tileinv, d_info, stream, false);
}

// Few right-hand sides: HBM-read bound (the lower triangle is read once per direction).  Blocks of VBLK
// columns: inside a block the 64-wide leaf steps (tile inverse + rank-64 update of the block's own rows),
// then ONE matrix-vector product with the whole panel below (forward) / left of (backward) the block, which
// streams >95 % of the factor at GEMV speed instead of in 782 latency-bound slivers.  Enqueue only.
// FVGP_TRSV_FUSED=0 keeps the step-per-launch substitutions at every size (A/B on the GPU box); read on every call.
static bool trsv_fused_enabled() {
  const char* e = getenv("FVGP_TRSV_FUSED");
  return e == nullptr || atoi(e) != 0;
}

// batch > 1: the same solve for `batch` problems; d_L, d_tileinv, d_B and d_work of problem b sit b * bstride doubles
// further (all inside equally sized workspace slots).
static int potrs_few(cudaStream_t st, const double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_B,
                     int nrhs, int64_t ldb, double* d_work, int batch = 1, long long bstride = 0) {
  const unsigned nb = (unsigned)batch;
  double* w = d_work;
  double* z = d_work + n;
  double* gemv_work = d_work + 2 * n;
  auto block_inv = [&](int t) {  // 64x64 diagonal sub-block of the 128x128 tile inverse (leading dimension TS)
    return d_tileinv + (int64_t)(t / 2) * TS * TS + (t % 2) * ((int64_t)VS * TS + VS);
  };
  const int nblocks = (int)((n + VBLK - 1) / VBLK);
  if (n <= TRSV_FUSED_MAX && trsv_fused_enabled()) {
    for (int r = 0; r < nrhs; ++r)
      launch(trsv_fused_kernel, nb, TRSV_FUSED_THREADS, 0, st, d_L, lda, (int)n, d_tileinv, w, z, d_B + (int64_t)r * ldb,
             bstride);
    FVGP_LAUNCH_OK();
    return 0;
  }
  for (int r = 0; r < nrhs; ++r) {
    double* b = d_B + (int64_t)r * ldb;
    if (batch == 1) {
      FVGP_CUDA_OK(cudaMemcpyAsync(w, b, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    } else {
      FVGP_CUDA_OK(cudaMemcpy2DAsync(w, bstride * sizeof(double), b, bstride * sizeof(double), n * sizeof(double), batch,
                                     cudaMemcpyDeviceToDevice, st));
    }
    // FVGP_TRSV_BLOCKS=1: walk a whole 2048-column block in ONE single-CTA launch (trsv_block_*_kernel).  Measured on
    // B200 it is SLOWER than one launch per 64-wide step (N = 8192: 5.13 vs 4.36 ms, N = 16 384: 11.4 vs 9.3 ms): the
    // step kernels spread the update of the remaining rows over many CTAs, which outweighs their launch latency.
    static int fused_env = -1;
    if (fused_env < 0) {
      const char* e = getenv("FVGP_TRSV_BLOCKS");
      fused_env = (e && atoi(e) == 1) ? 1 : 0;
    }
    const bool fused_blocks = fused_env == 1;
    for (int blk = 0; blk < nblocks; ++blk) {  // L z = b
      const int b0 = blk * VBLK, b1 = (int)std::min<int64_t>(n, b0 + VBLK);
      if (fused_blocks) {
        launch(trsv_block_fwd_kernel, nb, TRSV_FUSED_THREADS, 0, st, d_L, lda, b0, b1, d_tileinv, w, z, bstride);
      } else {
        for (int j0 = b0; j0 < b1; j0 += VS) {
          const int rest = b1 - (j0 + VS);
          const int grid = rest > 0 ? (rest + 63) / 64 : 1;
          launch(fwd_step_kernel, dim3(grid, nb), 256, 0, st, d_L, lda, b1, j0, block_inv(j0 / VS), w, z, bstride);
        }
      }
      if (b1 < n) {
        const int m = (int)n - b1;
        const int grid = (int)std::min<long long>(((long long)m + 7) / 8, (long long)sm_count() * 16);
        launch(gemv_n_kernel, dim3(grid, nb), 256, 0, st, d_L + (int64_t)b1 * lda + b0, lda, m, b1 - b0, -1.0, z + b0,
               w + b1, bstride);
      }
    }
    FVGP_LAUNCH_OK();
    for (int blk = nblocks - 1; blk >= 0; --blk) {  // L^T x = z
      const int b0 = blk * VBLK, b1 = (int)std::min<int64_t>(n, b0 + VBLK);
      const int last = b0 + ((b1 - b0 - 1) / VS) * VS;
      if (fused_blocks) {
        launch(trsv_block_bwd_kernel, nb, TRSV_FUSED_THREADS, 0, st, d_L, lda, b0, b1, d_tileinv, z, b, bstride);
      } else {
        for (int j0 = last; j0 >= b0; j0 -= VS) {
          const int cols = j0 - b0;
          const int grid = cols > 0 ? (cols + 255) / 256 : 1;
          launch(bwd_step_kernel, dim3(grid, nb), 256, 0, st, d_L, lda, b1, j0, block_inv(j0 / VS), z, b, b0, bstride);
        }
      }
      if (b0 > 0) {  // z[0:b0] -= L[b0:b1, 0:b0]^T x[b0:b1]
        const int m = b1 - b0;
        const int chunks = (m + GEMVT_ROWS - 1) / GEMVT_ROWS;
        launch(gemv_t_partial_kernel, dim3((b0 + 255) / 256, chunks, nb), 256, 0, st, d_L + (int64_t)b0 * lda, lda, m, b0,
               b + b0, gemv_work, bstride);
        launch(gemv_t_reduce_kernel, dim3((b0 + 255) / 256, nb), 256, 0, st, gemv_work, chunks, b0, -1.0, z, bstride);
      }
    }
    FVGP_LAUNCH_OK();
  }
  return 0;
}

int fvgp_potrs_lower(const double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_B, int nrhs,
                     int64_t ldb, double* d_work, void* stream) {
  FVGP_REQUIRE(n > 0 && nrhs >= 0 && ldb >= n && lda % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (nrhs > 4) {
    FVGP_REQUIRE(ldb % 2 == 0);
    Ctx c{st, const_cast<double*>(d_tileinv), nullptr, nullptr, 0};
    int r = trsm_rt_rec(c, d_B, ldb, nrhs, d_L, lda, (int)n, 0);   // rows of B <- rows * L^-T  (L y = b)
    if (r != 0) return r;
    return trsm_rn_rec(c, d_B, ldb, nrhs, d_L, lda, (int)n, 0);    // rows <- rows * L^-1      (L^T x = y)
  }
  return potrs_few(st, d_L, n, lda, d_tileinv, d_B, nrhs, ldb, d_work);
}

int fvgp_chol_logdet(const double* d_L, int64_t n, int64_t lda, double* d_scratch1, double* h_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  launch(logdet_kernel, 1, 1024, 0, st, d_L, lda, (int)n, d_scratch1, 0ll);
  FVGP_LAUNCH_OK();
  FVGP_CUDA_OK(cudaMemcpyAsync(h_out, d_scratch1, sizeof(double), cudaMemcpyDeviceToHost, st));
  FVGP_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int fvgp_dot(const double* d_a, const double* d_b, int64_t n, double* d_scratch1, double* h_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  launch(dot_kernel, 1, 1024, 0, st, d_a, d_b, n, d_scratch1);
  FVGP_LAUNCH_OK();
  FVGP_CUDA_OK(cudaMemcpyAsync(h_out, d_scratch1, sizeof(double), cudaMemcpyDeviceToHost, st));
  FVGP_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int fvgp_potri_lower(double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_work, void* stream) {
  FVGP_REQUIRE(n > 0 && n < (1ll << 31) && lda >= n && lda % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c{st, const_cast<double*>(d_tileinv), nullptr, d_work, 0};
  launch(zero_upper_diag_blocks_kernel, (unsigned)((n + BM - 1) / BM), 256, 0, st, d_L, lda, (int)n, 0ll);
  FVGP_LAUNCH_OK();
  // INT8-slice SYRK updates inside LAUUM (P11 += M21^T M21, half of LAUUM's flops), under the same switch as the POTRF
  // updates (fvgp_set_ozaki) and, like them, by default for N >= 40 000: POTRI at N = 50 000 2.479 -> 2.29 s, gradient
  // within 5.7e-10 of the DMMA path (LML unaffected), oracle parity at N = 16 000 with the threshold lowered
  // (profiles/r02/ozaki_lauum_step_probe.v13.log).  FVGP_OZAKI_LAUUM=0 switches it off, =<rows> sets the smallest P11
  // and lifts the N >= 40 000 gate (tests at smaller N).  FVGP_OZAKI_TRI=<chunks> (fvgp_set_ozaki_tri) additionally
  // sends the products with a triangular operand (both TRTRI products, W = M22^T M21 of LAUUM) through chunked INT8
  // GEMMs: trtri_products_int8 / lauum_w_int8 above.
  static int lauum_oz = -1;
  if (lauum_oz < 0) {
    const char* e = getenv("FVGP_OZAKI_LAUUM");
    lauum_oz = e ? atoi(e) : 1;
    if (lauum_oz > 1) g_ozaki_lauum_min = lauum_oz, g_ozaki_potri_min_n = 0;
  }
  const int n1_top = split((int)n), n2_top = (int)n - n1_top;
  if (lauum_oz > 0 && n >= g_ozaki_potri_min_n && ozaki_slices() > 0 && n > TS && n1_top >= g_ozaki_lauum_min) {
    const int S = ozaki_slices(), tri = ozaki_tri_chunks();
    const long long k16 = (n2_top + 15) / 16 * 16;
    const long long t_bytes = align256((long long)n1_top * k16 * 8);
    const long long syrk_bytes = t_bytes + fvgp_ozaki_work_bytes(n1_top, n1_top, k16, S, OZAKI_NBLOCK);
    c.oz_bytes = syrk_bytes;
    if (tri > 0) {
      const int cb1 = oz_chunk(n1_top, tri), cb2 = oz_chunk(n2_top, tri), cbm = std::max(cb1, cb2), nm = std::max(n1_top, n2_top);
      const long long trtri_need = std::max(align256((long long)n1_top * n1_top * 8), align256((long long)cb2 * n2_top * 8)) +
                                   fvgp_ozaki_work_bytes(cbm, nm, nm, S, OZAKI_NBLOCK);
      const long long w_need = t_bytes + align256((long long)n2_top * n2_top * 8) + fvgp_ozaki_work_bytes(cb2, n1_top, n2_top, S, OZAKI_NBLOCK);
      c.oz_bytes = std::max(c.oz_bytes, std::max(trtri_need, w_need));
    }
    pool_retention(n);
    // scratch ladder: everything, the SYRK half alone, then the SYRK with its contraction range in 2 / 4 / 8 pieces
    // (lauum_rec cuts it to whatever it is given; the triangular products refuse when their share does not fit)
    const long long full = c.oz_bytes;
    long long ladder[5] = {full, syrk_bytes, 0, 0, 0};
    for (int i = 2, parts = 2; i < 5; ++i, parts *= 2)
      ladder[i] = t_bytes + fvgp_ozaki_work_bytes(n1_top, n1_top, ((k16 + parts - 1) / parts + 15) / 16 * 16, S, OZAKI_NBLOCK);
    c.oz_work = nullptr;
    for (int i = 0; i < 5 && c.oz_work == nullptr; ++i) {
      if (i > 0 && ladder[i] >= ladder[i - 1]) continue;
      c.oz_bytes = ladder[i] + 4096;
      if (cudaMallocAsync(&c.oz_work, (size_t)c.oz_bytes, st) != cudaSuccess) {
        cudaGetLastError();
        c.oz_work = nullptr;
      }
    }
    if (c.oz_work != nullptr) {
      c.oz_slices = S;
      c.oz_tri = tri;
    } else {
      c.oz_bytes = 0;
    }
    static int told = -1;
    if (told != tri) {
      told = tri;
      fprintf(stderr, "[fvgp_b200] potri n=%lld: INT8-slice SYRK updates inside LAUUM%s, scratch %.2f GB: %s\n", (long long)n,
              c.oz_tri > 0 ? " + chunked triangular products" : "", c.oz_bytes / 1e9, c.oz_work ? "on" : "allocation failed, DMMA");
    }
  }
  int r = trtri_rec(c, d_L, lda, (int)n, 0);
  if (r != 0) {
    if (c.oz_work != nullptr) cudaFreeAsync(c.oz_work, st);
    return r;
  }
  r = lauum_rec(c, d_L, lda, (int)n);
  if (c.oz_work != nullptr) cudaFreeAsync(c.oz_work, st);
  return r;
}

int fvgp_dgemm_nt(const double* d_A, int64_t lda, const double* d_B, int64_t ldb, double* d_C, int64_t ldc, int m,
                  int n, int k, double alpha, double beta, int lower, void* stream) {
  return launch_gemm<false, false>((cudaStream_t)stream, d_A, lda, d_B, ldb, d_C, ldc, m, n, k, alpha, beta,
                                   lower ? GEMM_LOWER : 0);
}


/* ---- building blocks of the block-cyclic multi-GPU factorisation (fvgp_b200/sharded.py) ---- */

int fvgp_ozaki_slices(void) { return ozaki_slices(); }

// fvgp_dgemm's operand layouts on the INT8-slice path: operands stored k x rows are transposed into scratch (k padded
// to a multiple of 16 with zero columns) so that the product takes the NT form of fvgp_ozaki_gemm_nt.
int64_t fvgp_ozaki_gemm_work_bytes(int a_mn, int b_mn, int64_t m, int64_t n, int64_t k, int slices, int64_t nblock) {
  const int64_t k16 = (k + 15) / 16 * 16;
  return fvgp_ozaki_work_bytes(m, n, k16, slices, nblock) + (a_mn ? align256(m * k16 * 8) : 0) +
         (b_mn ? align256(n * k16 * 8) : 0);
}

int fvgp_ozaki_gemm(int a_mn, int b_mn, const double* d_A, int64_t lda, const double* d_B, int64_t ldb, double* d_C,
                    int64_t ldc, int64_t m, int64_t n, int64_t k, double sign, int zero_c, int slices, int64_t nblock,
                    void* d_work, int64_t work_bytes, void* stream) {
  FVGP_REQUIRE(m > 0 && n > 0 && k > 0 && m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31));
  FVGP_REQUIRE(work_bytes >= fvgp_ozaki_gemm_work_bytes(a_mn, b_mn, m, n, k, slices, nblock));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t k16 = (k + 15) / 16 * 16;
  if ((!a_mn || !b_mn) && k16 != k) return FVGP_ERR_ARG;  // a K-major operand cannot be padded in place
  char* w = (char*)d_work;
  const double *A = d_A, *B = d_B;
  int64_t la = lda, lb = ldb;
  if (a_mn) {  // stored k x m -> m x k16
    double* At = (double*)w;
    w += align256(m * k16 * 8);
    launch(transpose_pad_kernel, dim3((unsigned)((m + 31) / 32), (unsigned)((k16 + 31) / 32)), 256, 0, st, At, (long long)k16, d_A,
           (long long)lda, (int)k, (int)m, (int)k16, 0);
    A = At, la = k16;
  }
  if (b_mn) {  // stored k x n -> n x k16
    double* Bt = (double*)w;
    w += align256(n * k16 * 8);
    launch(transpose_pad_kernel, dim3((unsigned)((n + 31) / 32), (unsigned)((k16 + 31) / 32)), 256, 0, st, Bt, (long long)k16, d_B,
           (long long)ldb, (int)k, (int)n, (int)k16, 0);
    B = Bt, lb = k16;
  }
  FVGP_LAUNCH_OK();
  if (zero_c) FVGP_CUDA_OK(cudaMemset2DAsync(d_C, (size_t)ldc * sizeof(double), 0, (size_t)n * sizeof(double), (size_t)m, st));
  const int rc = fvgp_ozaki_gemm_nt(d_C, ldc, A, la, B, lb, m, n, k16, sign, 0, 0, 0, slices, nblock, w,
                                    work_bytes - (int64_t)(w - (char*)d_work), st);
  return rc;  // a refusal (< 0, not -100) after zero_c is still safe to redo with fvgp_dgemm and beta = 0
}

int fvgp_dgemm(int a_mn, int b_mn, const double* d_A, int64_t lda, const double* d_B, int64_t ldb, double* d_C,
               int64_t ldc, int m, int n, int k, double alpha, double beta, int flags, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!a_mn && !b_mn) return launch_gemm<false, false>(st, d_A, lda, d_B, ldb, d_C, ldc, m, n, k, alpha, beta, flags);
  if (!a_mn && b_mn) return launch_gemm<false, true>(st, d_A, lda, d_B, ldb, d_C, ldc, m, n, k, alpha, beta, flags);
  if (a_mn && b_mn) return launch_gemm<true, true>(st, d_A, lda, d_B, ldb, d_C, ldc, m, n, k, alpha, beta, flags);
  FVGP_REQUIRE(!"operand layout (A MN-major, B K-major) is not instantiated");
  return FVGP_ERR_ARG;
}

int fvgp_trsm_right_lower_t(double* d_B, int64_t ldb, int m, const double* d_L, int64_t ldl, int n,
                            const double* d_tileinv, void* stream) {
  FVGP_REQUIRE(n > 0 && m >= 0 && ldb % 2 == 0 && ldl % 2 == 0);
  Ctx c{(cudaStream_t)stream, const_cast<double*>(d_tileinv), nullptr, nullptr, 0};
  return trsm_rt_rec(c, d_B, ldb, m, d_L, ldl, n, 0);
}

int fvgp_trtri_lower(double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_work, void* stream) {
  FVGP_REQUIRE(n > 0 && n < (1ll << 31) && lda >= n && lda % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c{st, const_cast<double*>(d_tileinv), nullptr, d_work, 0};
  launch(zero_upper_diag_blocks_kernel, (unsigned)((n + BM - 1) / BM), 256, 0, st, d_L, lda, (int)n, 0ll);
  FVGP_LAUNCH_OK();
  return trtri_rec(c, d_L, lda, (int)n, 0);
}

int fvgp_lauum_lower(double* d_M, int64_t n, int64_t lda, double* d_work, void* stream) {
  FVGP_REQUIRE(n > 0 && n < (1ll << 31) && lda >= n && lda % 2 == 0);
  Ctx c{(cudaStream_t)stream, nullptr, nullptr, d_work, 0};
  return lauum_rec(c, d_M, lda, (int)n);
}

int fvgp_trsv_lower(const double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_b, int transpose,
                    double* d_work, void* stream) {
  FVGP_REQUIRE(n > 0 && lda % 2 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  double* w = d_work;
  double* z = d_work + n;
  const int tiles = (int)((n + VS - 1) / VS);
  auto block_inv = [&](int t) {
    return d_tileinv + (int64_t)(t / 2) * TS * TS + (t % 2) * ((int64_t)VS * TS + VS);
  };
  if (!transpose) {  // L z = b
    FVGP_CUDA_OK(cudaMemcpyAsync(w, d_b, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    for (int t = 0; t < tiles; ++t) {
      const int j0 = t * VS;
      const int rest = (int)n - (j0 + VS);
      const int grid = rest > 0 ? (rest + 63) / 64 : 1;
      launch(fwd_step_kernel, grid, 256, 0, st, d_L, lda, (int)n, j0, block_inv(t), w, z, 0ll);
    }
    FVGP_LAUNCH_OK();
    FVGP_CUDA_OK(cudaMemcpyAsync(d_b, z, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  } else {  // L^T x = b
    FVGP_CUDA_OK(cudaMemcpyAsync(z, d_b, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    for (int t = tiles - 1; t >= 0; --t) {
      const int j0 = t * VS;
      const int grid = j0 > 0 ? (j0 + 255) / 256 : 1;
      launch(bwd_step_kernel, grid, 256, 0, st, d_L, lda, (int)n, j0, block_inv(t), z, d_b, 0, 0ll);
    }
    FVGP_LAUNCH_OK();
  }
  return 0;
}

int64_t fvgp_gemv_work_len(int64_t m, int64_t n) { return ((m + GEMVT_ROWS - 1) / GEMVT_ROWS) * n + 1; }

int fvgp_gemv(int transpose, const double* d_A, int64_t lda, int m, int n, double alpha, const double* d_x,
              double* d_y, double* d_work, void* stream) {
  if (m <= 0 || n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (!transpose) {
    const int grid = (int)std::min<long long>(((long long)m + 7) / 8, (long long)sm_count() * 16);
    launch(gemv_n_kernel, grid, 256, 0, st, d_A, lda, m, n, alpha, d_x, d_y, 0ll);
  } else {
    const int chunks = (m + GEMVT_ROWS - 1) / GEMVT_ROWS;
    launch(gemv_t_partial_kernel, dim3((n + 255) / 256, chunks), 256, 0, st, d_A, lda, m, n, d_x, d_work, 0ll);
    launch(gemv_t_reduce_kernel, (n + 255) / 256, 256, 0, st, d_work, chunks, n, alpha, d_y, 0ll);
  }
  FVGP_LAUNCH_OK();
  return 0;
}

/* ---- population evaluation: many hyperparameter proposals per call ----
 * Two schedules behind one entry point:
 *   lock step (n < 6144, where one evaluation is a chain of ~70-700 launches that each fill a few SMs): the chain is
 *     launched ONCE for a whole chunk of proposals -- every kernel of the factorisation / solve / inverse takes the
 *     problem index from its grid (blockIdx.y / .z, the GEMM's BATCHED instantiation) and every buffer of problem b
 *     sits b * slot_len doubles further, so the host issues ~70 launches per chunk instead of per proposal;
 *   streams (larger n, where the look-ahead POTRF already fills the GPU): proposal b runs on stream b % slots. */

// slot layout (doubles): [A n*ld][tile inverses][potrs work][alpha 4n][res 16]( [potri work][trace partials] )
struct SlotLayout {
  int64_t ld, tileinv, swork, alpha, res, pwork, partials, len;
};
static SlotLayout slot_layout(int64_t n, int dim, int want_grad) {
  SlotLayout L;
  L.ld = round_up16(n);
  L.tileinv = round_up16(n * L.ld);
  L.swork = L.tileinv + round_up16(fvgp_chol_workspace_len(n));
  L.alpha = L.swork + round_up16(fvgp_potrs_work_len(n));
  L.res = L.alpha + round_up16(4 * n);
  L.pwork = L.res + 16;
  L.partials = L.pwork + (want_grad ? round_up16(fvgp_potri_workspace_len(n)) : 0);
  L.len = L.partials + (want_grad ? round_up16(fvgp_kgrad_partials_len(n, dim)) : 0);
  return L;
}

int64_t fvgp_population_slot_len(int64_t n, int dim, int want_grad) { return slot_layout(n, dim, want_grad).len; }

}  // extern "C"

namespace fvgp {
__global__ void broadcast_rows_kernel(double* dst, const double* __restrict__ src, long long count, long long bstride) {
  dst += blockIdx.y * bstride;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
}  // namespace fvgp

static double trace_fold(int kind, double length) {
  switch (kind) {
    case FVGP_K_MATERN32: return sqrt(3.0) / length;
    case FVGP_K_MATERN52: return sqrt(5.0) / length;
    case FVGP_K_SQEXP: return sqrt(0.5) / length;
    case FVGP_K_EXP: return 1.0 / length;
    default: return 0.0;
  }
}

// Lock-step schedule.  res_h: batch x (dim + 2) raw results [logdet, raw trace sums].
static int population_lockstep(int kind, const double* d_x, int64_t n, int dim, int batch, const double* h_amp,
                               const double* h_inv_scale, const double* h_length, const double* h_centre,
                               const double* d_noise, const double* d_rhs, int nrhs, int want_grad, int grad_component,
                               int slots, double* d_work, double* d_scratch, int64_t scratch_len, int* d_info,
                               double* h_alpha, double* res_h, cudaStream_t S) {
  const SlotLayout L = slot_layout(n, dim, want_grad);
  const int res_stride = dim + 2;
  for (int b0 = 0; b0 < batch; b0 += slots) {
    const int nb = std::min(slots, batch - b0);
    double* A0 = d_work;
    if (kfill_batch_theta_len(nb) <= scratch_len) {  // one fill launch for the whole chunk
      REC_OK(kfill_lower_batch_enqueue(kind, d_x, n, dim, nb, h_amp + b0, h_inv_scale + (int64_t)b0 * dim, h_length + b0,
                                       h_centre, d_noise, A0, L.ld, L.len, d_scratch, S));
    } else {
      for (int b = 0; b < nb; ++b) {
        int rc = fvgp_kfill_dense(kind, FVGP_FILL_LOWER, d_x, n, d_x, n, dim, h_amp[b0 + b],
                                  h_inv_scale + (int64_t)(b0 + b) * dim, h_centre, h_length[b0 + b], d_noise,
                                  A0 + (int64_t)b * L.len, L.ld, S);
        if (rc != 0) return rc;
      }
    }
    Ctx c{S, A0 + L.tileinv, d_info + b0, A0 + L.pwork, 0};
    c.batch = nb, c.bstride = L.len;
    REC_OK(potrf_rec(c, A0, L.ld, (int)n, 0));
    const long long cnt = (long long)nrhs * n;
    launch(broadcast_rows_kernel, dim3((unsigned)std::min<long long>((cnt + 255) / 256, 64), (unsigned)nb), 256, 0, S,
           A0 + L.alpha, d_rhs, cnt, (long long)L.len);
    REC_OK(potrs_few(S, A0, n, L.ld, A0 + L.tileinv, A0 + L.alpha, nrhs, n, A0 + L.swork, nb, L.len));
    launch(logdet_kernel, (unsigned)nb, 1024, 0, S, A0, L.ld, (int)n, A0 + L.res, (long long)L.len);
    FVGP_LAUNCH_OK();
    if (want_grad) {
      launch(zero_upper_diag_blocks_kernel, dim3((unsigned)((n + BM - 1) / BM), (unsigned)nb), 256, 0, S, A0, L.ld, (int)n,
             (long long)L.len);
      REC_OK(trtri_rec(c, A0, L.ld, (int)n, 0));
      REC_OK(lauum_rec(c, A0, L.ld, (int)n));
      for (int b = 0; b < nb; ++b) {
        const double fold = trace_fold(kind, h_length[b0 + b]);
        FVGP_REQUIRE(fold > 0.0);
        double coord[kMaxDim];
        for (int i = 0; i < dim; ++i) coord[i] = h_inv_scale[(int64_t)(b0 + b) * dim + i] * fold;
        double* slot = A0 + (int64_t)b * L.len;
        REC_OK(trace_radial_enqueue(kind, d_x, n, dim, coord, nullptr, slot, L.ld, slot + L.alpha + (int64_t)grad_component * n,
                                    slot + L.partials, slot + L.res + 1, S));
      }
    }
    FVGP_CUDA_OK(cudaMemcpy2DAsync(h_alpha + (int64_t)b0 * cnt, cnt * sizeof(double), A0 + L.alpha, L.len * sizeof(double),
                                   cnt * sizeof(double), nb, cudaMemcpyDeviceToHost, S));
    FVGP_CUDA_OK(cudaMemcpy2DAsync(res_h + (int64_t)b0 * res_stride, res_stride * sizeof(double), A0 + L.res,
                                   L.len * sizeof(double), res_stride * sizeof(double), nb, cudaMemcpyDeviceToHost, S));
  }
  return 0;
}

extern "C" {

int fvgp_lml_population(int kind, const double* d_x, int64_t n, int dim, int batch, const double* h_amp,
                        const double* h_inv_scale, const double* h_length, const double* h_centre,
                        const double* d_noise, const double* d_rhs, int nrhs, int want_grad, int grad_component,
                        int slots, double* d_work, double* d_alpha, double* d_res, int* d_info, double* h_alpha,
                        double* h_logdet, double* h_traces, int* h_info, void* stream) {
  FVGP_REQUIRE(n > 0 && n < (1ll << 31) && dim >= 1 && dim <= kMaxDim && batch >= 1 && slots >= 1);
  FVGP_REQUIRE(nrhs >= 1 && nrhs <= 4 && grad_component >= 0 && grad_component < nrhs);
  FVGP_REQUIRE(!want_grad || trace_fold(kind, 1.0) > 0.0);
  cudaStream_t S = (cudaStream_t)stream;
  const SlotLayout lay = slot_layout(n, dim, want_grad);
  const int64_t ld = lay.ld;
  const int res_stride = dim + 2;              // [logdet, raw trace sums R_0..R_dim]
  if (slots > batch) slots = batch;
  static bool configured = false;
  if (!configured) {
    FVGP_CUDA_OK(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)POTRF_TILE_SMEM));
    FVGP_CUDA_OK(cudaFuncSetAttribute(potrf_tile2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)POTRF_TILE2_SMEM));
    FVGP_CUDA_OK(cudaFuncSetAttribute(potrf_tile3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)POTRF_TILE3_SMEM));
    configured = true;
  }
  std::vector<double> res_h((size_t)batch * res_stride, 0.0);
  FVGP_CUDA_OK(cudaMemsetAsync(d_info, 0, sizeof(int) * batch, S));
  const int nb = potrf_block_width((int)n);
  // FVGP_POPULATION_STREAMS=1: stream schedule at every size (A/B on the GPU box, tests); read on every call
  const char* env_streams = getenv("FVGP_POPULATION_STREAMS");
  const bool force_streams = env_streams != nullptr && atoi(env_streams) != 0;
  int rc = 0;
  if (nb == 0 && !force_streams) {
    rc = population_lockstep(kind, d_x, n, dim, batch, h_amp, h_inv_scale, h_length, h_centre, d_noise, d_rhs, nrhs,
                             want_grad, grad_component, slots, d_work, d_alpha, (int64_t)batch * nrhs * n, d_info, h_alpha,
                             res_h.data(), S);
    if (rc != 0) {
      cudaStreamSynchronize(S);
      return rc;
    }
  } else {
    std::vector<cudaStream_t> st(slots);
    cudaEvent_t ev_start;
    FVGP_CUDA_OK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    FVGP_CUDA_OK(cudaEventRecord(ev_start, S));  // after the memset and everything the caller queued before
    for (int s = 0; s < slots; ++s) {
      FVGP_CUDA_OK(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
      FVGP_CUDA_OK(cudaStreamWaitEvent(st[s], ev_start, 0));
    }
    for (int b = 0; b < batch && rc == 0; ++b) {
      const int s = b % slots;
      cudaStream_t q = st[s];
      double* A = d_work + (int64_t)s * lay.len;
      double* tileinv = A + lay.tileinv;
      double* alpha = d_alpha + (int64_t)b * nrhs * n;
      double* res = d_res + (int64_t)b * res_stride;
      const double* inv = h_inv_scale + (int64_t)b * dim;
      // K(theta_b) + diag(V), lower triangle
      rc = fvgp_kfill_dense(kind, FVGP_FILL_LOWER, d_x, n, d_x, n, dim, h_amp[b], inv, h_centre, h_length[b], d_noise, A,
                            ld, q);
      if (rc != 0) break;
      Ctx c{q, tileinv, d_info + b, A + lay.pwork, 0};
      rc = nb > 0 ? potrf_lookahead(c, A, ld, (int)n, nb) : potrf_rec(c, A, ld, (int)n, 0);
      if (rc != 0) break;
      // alpha_b = KV^-1 (y - m), one row per right-hand side
      if (cudaMemcpyAsync(alpha, d_rhs, sizeof(double) * nrhs * n, cudaMemcpyDeviceToDevice, q) != cudaSuccess) {
        rc = FVGP_ERR_CUDA;
        break;
      }
      rc = potrs_few(q, A, n, ld, tileinv, alpha, nrhs, n, A + lay.swork);
      if (rc != 0) break;
      launch(logdet_kernel, 1, 1024, 0, q, A, ld, (int)n, res, 0ll);
      if (want_grad) {
        // lower(A) <- lower(KV^-1); raw sums R_0 = sum W f, R_i = sum W h q_i with W = KV^-1 - b b^T
        launch(zero_upper_diag_blocks_kernel, (unsigned)((n + BM - 1) / BM), 256, 0, q, A, ld, (int)n, 0ll);
        rc = trtri_rec(c, A, ld, (int)n, 0);
        if (rc == 0) rc = lauum_rec(c, A, ld, (int)n);
        if (rc != 0) break;
        const double fold = trace_fold(kind, h_length[b]);
        double coord[kMaxDim];
        for (int i = 0; i < dim; ++i) coord[i] = inv[i] * fold;
        rc = trace_radial_enqueue(kind, d_x, n, dim, coord, nullptr, A, ld, alpha + (int64_t)grad_component * n,
                                  A + lay.partials, res + 1, q);
        if (rc != 0) break;
      }
    }
    if (cudaGetLastError() != cudaSuccess && rc == 0) rc = FVGP_ERR_CUDA;
    // join: the caller's stream continues after every slot
    for (int s = 0; s < slots; ++s) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) {
        cudaEventRecord(e, st[s]);
        cudaStreamWaitEvent(S, e, 0);
        cudaEventDestroy(e);
      }
    }
    if (rc == 0) {
      if (cudaMemcpyAsync(h_alpha, d_alpha, sizeof(double) * batch * nrhs * n, cudaMemcpyDeviceToHost, S) != cudaSuccess ||
          cudaMemcpyAsync(res_h.data(), d_res, sizeof(double) * batch * res_stride, cudaMemcpyDeviceToHost, S) != cudaSuccess)
        rc = FVGP_ERR_CUDA;
    }
    cudaStreamSynchronize(S);
    for (int s = 0; s < slots; ++s) cudaStreamDestroy(st[s]);
    cudaEventDestroy(ev_start);
    if (rc != 0) return rc;
  }
  FVGP_CUDA_OK(cudaMemcpyAsync(h_info, d_info, sizeof(int) * batch, cudaMemcpyDeviceToHost, S));
  FVGP_CUDA_OK(cudaStreamSynchronize(S));
  for (int b = 0; b < batch; ++b) {
    h_logdet[b] = res_h[(size_t)b * res_stride];
    if (want_grad) {  // descriptor traces (T_amp, T_s1..T_sD, T_length), as fvgp_kgrad_trace_radial
      const double* raw = &res_h[(size_t)b * res_stride + 1];
      double* out = h_traces + (size_t)b * (dim + 2);
      double sum = 0.0;
      out[0] = raw[0];
      for (int i = 0; i < dim; ++i) {
        const double si = h_inv_scale[(size_t)b * dim + i];
        sum += raw[1 + i];
        out[1 + i] = si != 0.0 ? -(h_amp[b] / si) * raw[1 + i] : 0.0;
      }
      out[dim + 1] = (h_amp[b] / h_length[b]) * sum;
    }
  }
  return 0;
}

}  // extern "C"
