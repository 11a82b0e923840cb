// Dense covariance assembly and gradient traces.
//
//   fvgp_kfill_dense           K = amp * f(||(x1-x2)*inv_scale|| / length) (+ noise on the diagonal)
//                              replaces gp_prior.py:376-400, kernels.py:16-188, :440-481, gp_kv.py:640-669
//   fvgp_kgrad_trace_matern32  sum_ij (Kinv - b b^T)_ij dK_ij/dtheta_h with dK regenerated per tile
//                              replaces gp_prior.py:421-436 + gp_marginal_likelihood.py:264-306
//   fvgp_kgrad_dense_matern32  materialised dK/dtheta (kernel_function_grad seam)
//
// K-fill is HBM-write bound (8 bytes per entry) with the FP64 pipe a close second
// (~40 DP instructions per entry: sqrt + exp).  The symmetric mode therefore evaluates
// only the upper tile triangle: a 64x64 tile is stored directly from registers (each
// warp store covers 256 contiguous bytes of one row) and, transposed through padded
// shared memory, a second time as the mirror tile with one TMA bulk store
// (cp.async.bulk.global.shared::cta) per 512-byte row, double buffered so the stores of
// tile t drain under the arithmetic of tile t+1.  Differences are taken directly
// (x1-x2), never through the |x|^2+|y|^2-2xy expansion: the expansion loses the 1e-12
// relative entry accuracy once coordinates exceed ~40 length scales (hard part 3).
#include "../../include/fvgp_b200.h"
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace fvgp {

constexpr int FT = 64;             // tile edge
constexpr int FT_STRIDE = FT + 2;  // even (16-byte rows for the bulk copy)
// Staging row c of the transposed tile starts stage_shift(c) doubles into its 66-double slot.  The transposed
// stores are 16-byte stores to rows 2l / 2l+1 (l = lane): a quarter warp (8 lanes, what one 128-bit shared-memory
// wavefront serves) then walks 8 rows whose starts are 2*66*8 bytes apart = 8 banks apart, so lanes l and l+4 of
// every quarter met in the same 4 banks (2-way conflict, 85.7 M conflicts per N = 50 000 fill in ncu r01).  Shifting
// rows 8..15, 24..31, ... by 16 bytes moves the second half of each quarter onto the 4 banks in between: conflict
// free, rows stay 16-byte aligned and contiguous for the bulk copy, and 64 + 2 doubles still fit the slot.
__device__ __forceinline__ int stage_shift(int c) { return 2 * ((c >> 3) & 1); }
constexpr int FILL_THREADS = 256;
constexpr int STAGE_DOUBLES = FT * FT_STRIDE;
// Per-warp mirror staging (bulk mode 2): every warp transposes its own 8 x 64 strip into a 64 x 8 block (row slot of
// 10 doubles = 8 values + the 16-byte bank shift) and sends it with 64-byte bulk stores of its own -- no CTA-wide
// barrier anywhere in the tile loop, the warps of a CTA drift apart and hide each other's sqrt / exp chains.
constexpr int WSTRIDE = 10;
constexpr int WSTAGE_DOUBLES = FT * WSTRIDE;  // per warp

static int g_use_bulk_store = 1;  // 2 (per-warp 64-byte bulk stores) measured SLOWER on B200: 6.22 vs 4.34 ms at N = 50 000

struct FillParams {
  const double* x1;
  const double* x2;
  const double* noise;
  double* K;
  long long n1, n2, ldk;
  double amp, c_arg, c_aux;  // kind-specific constants
  double inv_scale[kMaxDim];  // centred path: inv_scale * (kind's argument constant)
  double centre[kMaxDim];
  int dim, mode;
  long long tiles_i, tiles_j, ntiles;
  int bulk, vec2, band;
};

__device__ __forceinline__ void tri_index(long long id, long long& a, long long& b) {
  a = (long long)((sqrt(8.0 * (double)id + 1.0) - 1.0) * 0.5);
  while (a * (a + 1) / 2 > id) --a;
  while ((a + 1) * (a + 2) / 2 <= id) ++a;
  b = id - a * (a + 1) / 2;
}

// Lower-triangular tile enumeration in BANDS of G tile rows, column by column inside a band (the order the
// GEMM rasteriser uses).  A CTA that walks a contiguous id range then stays inside G*64 matrix rows for its
// mirror stores and advances its direct stores by one 64-row block per G tiles, instead of sweeping a whole
// matrix column: at N = 50 000 (400 KB row pitch) the column sweep touched ~26 distinct 2 MB pages per tile
// and the symmetric fill dropped from 4.6 TB/s (N = 30 000) to 2.9 TB/s.
__device__ __forceinline__ void band_index(long long id, long long tiles, int G, long long& a, long long& b) {
  // tiles in bands 0..k-1: G*G*k*(k-1)/2 + k*G*(G+1)/2
  const double g = (double)G;
  long long k = (long long)((sqrt(8.0 * (double)id / (g * g) + 1.0) - 1.0) * 0.5);
  auto before = [&](long long kk) { return (long long)G * G * kk * (kk - 1) / 2 + kk * (long long)G * (G + 1) / 2; };
  while (k > 0 && before(k) > id) --k;
  while (before(k + 1) <= id && (k + 1) * G < tiles) ++k;
  const long long r0 = k * G;
  const long long gh = min((long long)G, tiles - r0);
  // a ragged last band has fewer rows: its offset formula is still before(k) (all earlier bands are full)
  long long l = id - before(k);
  if (l < r0 * gh) {
    b = l / gh;
    a = r0 + l - b * gh;
    return;
  }
  l -= r0 * gh;
  long long c = 0;
  while (l >= gh - c) {
    l -= gh - c;
    ++c;
  }
  b = r0 + c;
  a = r0 + c + l;
}

// ---- lean FP64 math for the fill.  The kernel is bound by the FP64 pipe AND by instruction issue
// (ncu: 37 DP + 37 other instructions per entry with libdevice-free but naive code), so every
// special case is moved to the integer pipe and every constant comes from the constant bank:
//   exp(-a), a >= 0 : Cody-Waite reduction + degree-11 Horner; the 2^n scaling is an integer add on
//                     the high word with n clamped at -1022, so results below 2.3e-308 come back
//                     as (sub)normal garbage of at most that magnitude instead of exactly 0;
//   sqrt(s), s >= 0 : MUFU.RSQ64H on the high word (clamped to the smallest normal with an integer
//                     max, which also makes s = 0 return exactly 0) + one Newton step + residual
//                     correction, ~1 ulp.
// Tests compare K entries with the reference at 1e-12 relative.
__constant__ double kExpC[12] = {2.5022322536502990E-008, 2.7630903488173108E-007, 2.7557514545882439E-006,
                                 2.4801491039099165E-005, 1.9841269589115497E-004, 1.3888888945916380E-003,
                                 8.3333333334550432E-003, 4.1666666666519754E-002, 1.6666666666666477E-001,
                                 5.0000000000000122E-001, 1.0, 1.0};
__constant__ double kExpK[4] = {-1.4426950408889634, 6755399441055744.0, -6.93147180559945286e-01,
                                -2.31904681384629956e-17};

__device__ __forceinline__ double exp_neg(double a) {  // exp(-a) for 0 <= a < 2^30
  const double t = fma(a, kExpK[0], kExpK[1]);
  int n = __double2loint(t);
  const double nf = t - kExpK[1];
  double r = fma(nf, kExpK[2], -a);
  r = fma(nf, kExpK[3], r);
  double p = kExpC[0];
#pragma unroll
  for (int i = 1; i < 12; ++i) p = fma(p, r, kExpC[i]);
  n = max(n, -1022);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

__device__ __forceinline__ double exp_nonpos(double x) {  // x <= 0 (trace kernel, elementwise kernels)
  return exp_neg(fmin(-x, 1.0e9));
}

__device__ __forceinline__ double sqrt_pos(double s) {  // s >= 2.3e-308 (normal), finite
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double g = s * y;
  const double h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  const double d = fma(-g, g, s);
  return fma(d, h, g);
}

// Value from the squared scaled distance s.  CENTRED: the kind's argument constant is already folded
// into the coordinate scaling, so that a = sqrt(s) (or s) is the argument of exp directly.
template <int KIND, bool CENTRED>
__device__ __forceinline__ double radial_value(double s, double amp, double c_arg, double c_aux) {
  if (KIND == FVGP_K_SQEXP) return amp * exp_neg(CENTRED ? s : fmin(s * c_arg, 1.0e9));  // c_arg = 1/(2 l^2)
  const double d = ((KIND == FVGP_K_WENDLAND && !CENTRED) || KIND == FVGP_K_DISTANCE) ? sqrt(s) : sqrt_pos(s);
  if (KIND == FVGP_K_DISTANCE) return d;
  if (KIND == FVGP_K_MATERN32) {
    const double a = CENTRED ? d : fmin(c_arg * d, 1.0e9);  // c_arg = sqrt(3)/l
    return fma(amp, a, amp) * exp_neg(a);
  }
  if (KIND == FVGP_K_MATERN52) {
    const double a = CENTRED ? d : fmin(c_arg * d, 1.0e9);  // c_arg = sqrt(5)/l, c_aux = amp * 5/(3 l^2) | amp/3
    return fma(c_aux, s, fma(amp, a, amp)) * exp_neg(a);
  }
  if (KIND == FVGP_K_EXP) return amp * exp_neg(CENTRED ? d : fmin(d * c_arg, 1.0e9));  // c_arg = 1/l
  // Wendland (dense form, kernels.py:355-378)
  const double dd = fmin(CENTRED ? d : d * c_arg, 1.0);
  const double u = 1.0 - dd;
  const double u2 = u * u, u4 = u2 * u2;
  return amp * (u4 * u4) * (32.0 * dd * dd * dd + 25.0 * dd * dd + 8.0 * dd + 1.0);
}

__device__ __forceinline__ void bulk_store_row(double* gdst, const double* ssrc, unsigned bytes) {
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(ssrc));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes)
               : "memory");
}

// One 64x64 tile.  Warp w owns rows 8w..8w+7; lane l owns the ADJACENT columns 2l, 2l+1, so a warp stores one
// full 512-byte tile row with a single STG.128 per lane.  Rows are processed two at a time: four independent
// sqrt/exp dependency chains per thread keep the FP64 pipe fed.  The warp's row coordinates are staged once per
// tile in shared memory (already centred and scaled in the CENTRED variant): the inner loop then has no global
// loads and, per entry, D subtractions + D fused multiply-adds for the distance.
// INTERIOR tiles carry no bounds checks and no noise test.
template <int KIND, int DIM, bool CENTRED, bool INTERIOR, bool WARP_STAGE = false>
__device__ __forceinline__ void fill_tile(const FillParams& p, long long ti, long long tj, bool mirror, double* sT,
                                          double* sRow, int lane, int warp) {
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  constexpr int DPAD = (D + 1) & ~1;
  const int dim = DIM > 0 ? DIM : p.dim;
  const long long r0 = ti * FT + warp * 8, c0 = tj * FT;
  const long long ca = c0 + 2 * lane, cb = ca + 1;
  double xa[D], xb[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const bool on = DIM > 0 || i < dim;
    xa[i] = ((INTERIOR || ca < p.n2) && on) ? p.x2[ca * dim + i] : 0.0;
    xb[i] = ((INTERIOR || cb < p.n2) && on) ? p.x2[cb * dim + i] : 0.0;
    if (CENTRED) {
      // intrinsics, not operators: the compiler must not fold this multiply into the subtractions of the inner
      // loop (x0 - m * s -> fma(-m, s, x0) skips the rounding of the scaled coordinate; the staged ROW
      // coordinates below are rounded, so a point would act with two slightly different positions and the
      // full and mirrored fills would disagree in the last bit)
      xa[i] = __dmul_rn(__dsub_rn(xa[i], p.centre[i]), p.inv_scale[i]);
      xb[i] = __dmul_rn(__dsub_rn(xb[i], p.centre[i]), p.inv_scale[i]);
    }
  }
  __syncwarp();  // the previous tile's reads of sRow are done
  for (int idx = lane; idx < 8 * dim; idx += 32) {
    const int rr = idx / dim, i = idx - rr * dim;
    double v = (INTERIOR || r0 + rr < p.n1) ? p.x1[(r0 + rr) * dim + i] : 0.0;
    if (CENTRED) v = __dmul_rn(__dsub_rn(v, p.centre[i]), p.inv_scale[i]);
    sRow[rr * DPAD + i] = v;
  }
  __syncwarp();
  double* krow = p.K + r0 * p.ldk + ca;
  constexpr int SSTR = WARP_STAGE ? WSTRIDE : FT_STRIDE;
  double* srow = sT + (2 * lane) * SSTR + stage_shift(2 * lane) + (WARP_STAGE ? 0 : warp * 8);
#pragma unroll 1
  for (int rr = 0; rr < 8; rr += 2) {
    if (!INTERIOR && r0 + rr >= p.n1) break;
    // 1e-300 instead of 0: keeps the reciprocal square root finite for coincident points at no cost (the
    // distance kind uses the IEEE square root and returns an exact 0)
    constexpr double S0 = KIND == FVGP_K_DISTANCE ? 0.0 : 1e-300;
    double s00 = S0, s01 = S0, s10 = S0, s11 = S0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (DIM > 0 || i < dim) {
        const double x0 = sRow[rr * DPAD + i], x1 = sRow[(rr + 1) * DPAD + i];
        double t00 = x0 - xa[i], t01 = x0 - xb[i], t10 = x1 - xa[i], t11 = x1 - xb[i];
        if (!CENTRED) {
          const double sc = p.inv_scale[i];
          t00 *= sc, t01 *= sc, t10 *= sc, t11 *= sc;
        }
        s00 = fma(t00, t00, s00), s01 = fma(t01, t01, s01), s10 = fma(t10, t10, s10), s11 = fma(t11, t11, s11);
      }
    }
    double v00 = radial_value<KIND, CENTRED>(s00, p.amp, p.c_arg, p.c_aux);
    double v01 = radial_value<KIND, CENTRED>(s01, p.amp, p.c_arg, p.c_aux);
    double v10 = radial_value<KIND, CENTRED>(s10, p.amp, p.c_arg, p.c_aux);
    double v11 = radial_value<KIND, CENTRED>(s11, p.amp, p.c_arg, p.c_aux);
    if (INTERIOR) {
      if (p.vec2) {
        *reinterpret_cast<double2*>(krow + rr * p.ldk) = make_double2(v00, v01);
        *reinterpret_cast<double2*>(krow + (rr + 1) * p.ldk) = make_double2(v10, v11);
      } else {
        krow[rr * p.ldk] = v00, krow[rr * p.ldk + 1] = v01;
        krow[(rr + 1) * p.ldk] = v10, krow[(rr + 1) * p.ldk + 1] = v11;
      }
    } else {
      const long long ra = r0 + rr, rb = ra + 1;
      if (p.noise != nullptr) {
        if (ra == ca) v00 += p.noise[ca];
        if (ra == cb) v01 += p.noise[cb];
        if (rb == ca) v10 += p.noise[ca];
        if (rb == cb) v11 += p.noise[cb];
      }
      if (ca < p.n2) krow[rr * p.ldk] = v00;
      if (cb < p.n2) krow[rr * p.ldk + 1] = v01;
      if (rb < p.n1) {
        if (ca < p.n2) krow[(rr + 1) * p.ldk] = v10;
        if (cb < p.n2) krow[(rr + 1) * p.ldk + 1] = v11;
      }
    }
    if (mirror) {  // transposed staging tile: sT[column][row]
      *reinterpret_cast<double2*>(srow + rr) = make_double2(v00, v10);
      *reinterpret_cast<double2*>(srow + SSTR + rr) = make_double2(v01, v11);
    }
  }
}

template <int KIND, int DIM, bool CENTRED>
__global__ void __launch_bounds__(FILL_THREADS, 3) kfill_kernel(const FillParams p) {
  // [8 warps][8 rows][DPAD] staged row coordinates, then two FT x FT_STRIDE staging tiles (symmetric mode only)
  extern __shared__ __align__(128) double fill_smem[];
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  constexpr int DPAD = (D + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* sRow = fill_smem + warp * 8 * DPAD;
  double* stage = fill_smem + 8 * 8 * DPAD;
  int buf = 0;

  // Each CTA walks a CONTIGUOUS range of tiles: the (row block, column block) pair is decoded once (one FP64
  // square root) and then advanced incrementally; consecutive tiles share a row / column block, so their
  // coordinates stay in L1.
  const long long per_cta = (p.ntiles + gridDim.x - 1) / gridDim.x;
  long long tile = blockIdx.x * per_cta;
  const long long tile_end = min(p.ntiles, tile + per_cta);
  long long ia = 0, ib = 0;
  const bool banded = p.band > 0 && p.mode == FVGP_FILL_SYMMETRIC;
  if (tile < tile_end && !banded) {
    if (p.mode == FVGP_FILL_FULL) {
      ia = tile / p.tiles_j;
      ib = tile - ia * p.tiles_j;
    } else {
      tri_index(tile, ia, ib);
    }
  }
  for (; tile < tile_end; ++tile) {
    if (banded) band_index(tile, p.tiles_i, p.band, ia, ib);
    long long ti = ia, tj = ib;
    if (p.mode == FVGP_FILL_SYMMETRIC) ti = ib, tj = ia;
    if (banded) {
    } else if (p.mode == FVGP_FILL_FULL) {
      if (++ib == p.tiles_j) ib = 0, ++ia;
    } else if (++ib > ia) {
      ib = 0, ++ia;
    }
    const bool mirror = (p.mode == FVGP_FILL_SYMMETRIC) && (tj > ti);
    if (p.bulk == 2) {  // per-warp mirror: no CTA barrier in the loop
      double* wT = stage + warp * WSTAGE_DOUBLES;
      if (mirror) {  // this warp's previous bulk stores must have finished READING its staging block
        asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        __syncwarp();
      }
      const bool interior2 = (ti + 1) * FT <= p.n1 && (tj + 1) * FT <= p.n2 && (p.noise == nullptr || ti != tj);
      if (interior2) fill_tile<KIND, DIM, CENTRED, true, true>(p, ti, tj, mirror, wT, sRow, lane, warp);
      else fill_tile<KIND, DIM, CENTRED, false, true>(p, ti, tj, mirror, wT, sRow, lane, warp);
      if (mirror) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        const long long c0 = tj * FT;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = lane + 32 * h;
          if (c0 + c < p.n2)
            bulk_store_row(p.K + (c0 + c) * p.ldk + ti * FT + warp * 8, wT + c * WSTRIDE + stage_shift(c), 8 * 8);
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
      }
      continue;
    }
    double* sT = stage + buf * STAGE_DOUBLES;
    if (mirror) {
      // the bulk stores that last used THIS staging buffer (two mirror tiles ago) must have finished reading it
      if (p.bulk && tid < FT) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
      __syncthreads();
    }
    const bool interior = (ti + 1) * FT <= p.n1 && (tj + 1) * FT <= p.n2 && (p.noise == nullptr || ti != tj);
    if (interior) fill_tile<KIND, DIM, CENTRED, true>(p, ti, tj, mirror, sT, sRow, lane, warp);
    else fill_tile<KIND, DIM, CENTRED, false>(p, ti, tj, mirror, sT, sRow, lane, warp);
    if (mirror) {
      // mirror tile: rows tj*FT.. of K, columns ti*FT .. ti*FT+63 (always a full 64 because ti < tj)
      const long long c0 = tj * FT;
      if (p.bulk) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (tid < FT) {
          const long long grow = c0 + tid;
          if (grow < p.n2) bulk_store_row(p.K + grow * p.ldk + ti * FT, sT + tid * FT_STRIDE + stage_shift(tid), FT * 8);
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        buf ^= 1;
      } else {
        __syncthreads();
#pragma unroll 2
        for (int rr = 0; rr < 8; ++rr) {
          const long long grow = c0 + warp * 8 + rr;
          if (grow >= p.n2) break;
          double* krow = p.K + grow * p.ldk + ti * FT;
          const double* src = sT + (warp * 8 + rr) * FT_STRIDE + stage_shift(warp * 8 + rr);
          krow[lane] = src[lane];
          krow[lane + 32] = src[lane + 32];
        }
      }
    }
  }
  if (p.bulk == 2) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
  else if (p.bulk && tid < FT) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}

template <int KIND, int DIM, bool CENTRED>
static void launch_fill_one(const FillParams& p, unsigned grid, cudaStream_t st) {
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  constexpr int DPAD = (D + 1) & ~1;
  const size_t stage = p.mode != FVGP_FILL_SYMMETRIC ? 0 : (p.bulk == 2 ? 8 * WSTAGE_DOUBLES : 2 * STAGE_DOUBLES);
  const size_t smem = (8 * 8 * DPAD + stage) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(kfill_kernel<KIND, DIM, CENTRED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((8 * 8 * DPAD + 2 * STAGE_DOUBLES) * sizeof(double)));
    configured = true;
  }
  launch(kfill_kernel<KIND, DIM, CENTRED>, grid, FILL_THREADS, smem, st, p);
}

template <int KIND>
static int launch_fill_dim(const FillParams& p, bool centred, cudaStream_t st) {
  const long long max_ctas = (long long)sm_count() * 3;
  const unsigned grid = (unsigned)(p.ntiles < max_ctas ? p.ntiles : max_ctas);
#define FVGP_FILL_CASE(DIMV)                                              \
  {                                                                       \
    if (centred) launch_fill_one<KIND, DIMV, true>(p, grid, st);          \
    else launch_fill_one<KIND, DIMV, false>(p, grid, st);                 \
  }
  switch (p.dim) {
    case 1: FVGP_FILL_CASE(1) break;
    case 2: FVGP_FILL_CASE(2) break;
    case 3: FVGP_FILL_CASE(3) break;
    case 4: FVGP_FILL_CASE(4) break;
    default: launch_fill_one<KIND, 0, false>(p, grid, st); break;
  }
#undef FVGP_FILL_CASE
  FVGP_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Batched fill for the population evaluator (dense_linalg.cu, lock-step schedule): blockIdx.y selects a proposal,
// whose (amp, c_arg, c_aux, inv_scale) come from a device array and whose matrix sits blockIdx.y * kstride doubles
// after the first one.  Same tile routine (fill_tile) and constants as the single fill -> bitwise the same matrices.
// FULL / LOWER modes only (the factorisation wants the lower triangle; no mirror staging).
// ----------------------------------------------------------------------------------------------
struct FillTheta {
  double amp, c_arg, c_aux;
  double inv_scale[kMaxDim];
};

template <int KIND, int DIM, bool CENTRED>
__global__ void __launch_bounds__(FILL_THREADS, 3) kfill_batch_kernel(const FillParams base,
                                                                      const FillTheta* __restrict__ thetas,
                                                                      long long kstride) {
  extern __shared__ __align__(128) double fill_smem[];
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  constexpr int DPAD = (D + 1) & ~1;
  FillParams p = base;
  {
    const FillTheta* t = thetas + blockIdx.y;
    p.amp = t->amp, p.c_arg = t->c_arg, p.c_aux = t->c_aux;
#pragma unroll
    for (int i = 0; i < kMaxDim; ++i) p.inv_scale[i] = t->inv_scale[i];
    p.K = base.K + (long long)blockIdx.y * kstride;
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* sRow = fill_smem + warp * 8 * DPAD;
  const long long per_cta = (p.ntiles + gridDim.x - 1) / gridDim.x;
  long long tile = blockIdx.x * per_cta;
  const long long tile_end = min(p.ntiles, tile + per_cta);
  long long ia = 0, ib = 0;
  if (tile < tile_end) {
    if (p.mode == FVGP_FILL_FULL) {
      ia = tile / p.tiles_j;
      ib = tile - ia * p.tiles_j;
    } else {
      tri_index(tile, ia, ib);
    }
  }
  for (; tile < tile_end; ++tile) {
    const long long ti = ia, tj = ib;
    if (p.mode == FVGP_FILL_FULL) {
      if (++ib == p.tiles_j) ib = 0, ++ia;
    } else if (++ib > ia) {
      ib = 0, ++ia;
    }
    const bool interior = (ti + 1) * FT <= p.n1 && (tj + 1) * FT <= p.n2 && (p.noise == nullptr || ti != tj);
    if (interior) fill_tile<KIND, DIM, CENTRED, true>(p, ti, tj, false, fill_smem, sRow, lane, warp);
    else fill_tile<KIND, DIM, CENTRED, false>(p, ti, tj, false, fill_smem, sRow, lane, warp);
  }
}

template <int KIND, int DIM, bool CENTRED>
static void launch_fill_batch_one(const FillParams& p, const FillTheta* d_thetas, long long kstride, int batch,
                                  cudaStream_t st) {
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  constexpr int DPAD = (D + 1) & ~1;
  const long long per_problem = std::max<long long>(1, ((long long)sm_count() * 3 + batch - 1) / batch);
  const unsigned gx = (unsigned)std::min<long long>(p.ntiles, per_problem);
  launch(kfill_batch_kernel<KIND, DIM, CENTRED>, dim3(gx, (unsigned)batch), FILL_THREADS, 8 * 8 * DPAD * sizeof(double), st,
         p, d_thetas, kstride);
}

template <int KIND>
static int launch_fill_batch_dim(const FillParams& p, bool centred, const FillTheta* d_thetas, long long kstride,
                                 int batch, cudaStream_t st) {
#define FVGP_FILL_BATCH_CASE(DIMV)                                                              \
  {                                                                                             \
    if (centred) launch_fill_batch_one<KIND, DIMV, true>(p, d_thetas, kstride, batch, st);      \
    else launch_fill_batch_one<KIND, DIMV, false>(p, d_thetas, kstride, batch, st);             \
  }
  switch (p.dim) {
    case 1: FVGP_FILL_BATCH_CASE(1) break;
    case 2: FVGP_FILL_BATCH_CASE(2) break;
    case 3: FVGP_FILL_BATCH_CASE(3) break;
    case 4: FVGP_FILL_BATCH_CASE(4) break;
    default: launch_fill_batch_one<KIND, 0, false>(p, d_thetas, kstride, batch, st); break;
  }
#undef FVGP_FILL_BATCH_CASE
  FVGP_LAUNCH_OK();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Gradient traces for the default ARD Matern-3/2 kernel.
// ----------------------------------------------------------------------------------------------
struct TraceParams {
  const double* x;
  const double* Kinv;
  const double* b;
  double* partials;
  long long n, ld, ntiles;
  double inv_len[kMaxDim];
  int dim;
};

// One 64x64 tile of the trace, same thread layout as fill_tile: warp w owns rows 8w..8w+7, lane l the adjacent
// columns 2l, 2l+1 (one 16-byte load of W per lane and row), rows two at a time -> four independent sqrt / exp
// chains per thread.  Row coordinates (already divided by the length scales) and b are staged per warp in shared
// memory.  INTERIOR tiles (strictly below the diagonal, fully inside the matrix) carry no masks: weight 2.
// KIND selects the radial family; the coordinates arrive scaled by inv_scale_i * c_KIND / length so that the sum of
// squared differences is a^2 (Matern-3/2: a = sqrt(3) u, Matern-5/2: sqrt(5) u, exponential: u) or u^2 / 2
// (squared exponential), u = r / length.  Accumulated per entry, with W = (KV^-1 - b b^T) weighted 2 / 1 / 0:
//   acc[0]   += W f(u)                 -> trace against dK/d(amp) (times amp^0)
//   acc[1+i] += W h(u) q_i             -> -s_i / amp times the trace against dK/d(inv_scale_i)
// with q_i the squared scaled difference along axis i and h = -g q-scaling (see fvgp_kgrad_trace_radial).
template <int KIND>
__device__ __forceinline__ void radial_trace_factors(double s, double& f, double& h) {
  if (KIND == FVGP_K_SQEXP) {
    const double e = exp_neg(fmin(s, 1.0e9));
    f = e, h = 2.0 * e;
    return;
  }
  const double a = sqrt_pos(s);
  const double e = exp_neg(fmin(a, 1.0e9));
  if (KIND == FVGP_K_MATERN32) f = (1.0 + a) * e, h = e;
  else if (KIND == FVGP_K_MATERN52) f = fma(s, 1.0 / 3.0, 1.0 + a) * e, h = (1.0 + a) * e * (1.0 / 3.0);
  else f = e, h = e / a;  // exponential; s >= 1e-300 keeps a > 0 and q_i = 0 for coincident points
}

template <int KIND, int DIM, bool INTERIOR>
__device__ __forceinline__ void trace_tile(const TraceParams& p, long long ti, long long tj, double* sRow,
                                           const double (&inv)[DIM > 0 ? DIM : kMaxDim], double (&acc)[(DIM > 0 ? DIM : kMaxDim) + 1],
                                           int lane, int warp) {
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  constexpr int DP = D + 1;  // staged per row: D scaled coordinates + b
  const int dim = DIM > 0 ? DIM : p.dim;
  const long long r0 = ti * FT + warp * 8, c0 = tj * FT;
  const long long ca = c0 + 2 * lane, cb = ca + 1;
  double xa[D], xb[D];
  const bool oka = INTERIOR || ca < p.n, okb = INTERIOR || cb < p.n;
  const double ba = oka ? p.b[ca] : 0.0, bb = okb ? p.b[cb] : 0.0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const bool on = DIM > 0 || i < dim;
    xa[i] = (oka && on) ? p.x[ca * dim + i] * inv[i] : 0.0;
    xb[i] = (okb && on) ? p.x[cb * dim + i] * inv[i] : 0.0;
  }
  __syncwarp();  // the previous tile's reads of sRow are done
  for (int idx = lane; idx < 8 * (dim + 1); idx += 32) {
    const int rr = idx / (dim + 1), i = idx - rr * (dim + 1);
    const long long r = r0 + rr;
    double v = 0.0;
    if (INTERIOR || r < p.n) v = (i < dim) ? p.x[r * dim + i] * p.inv_len[i] : p.b[r];
    sRow[rr * DP + (i < dim ? i : D)] = v;
  }
  __syncwarp();
  const bool vec = (p.ld % 2 == 0);
#pragma unroll 1
  for (int rr = 0; rr < 8; rr += 2) {
    const long long ra = r0 + rr, rb = ra + 1;
    if (!INTERIOR && ra >= p.n) break;
    const bool okr = INTERIOR || rb < p.n;
    const double* wrow = p.Kinv + ra * p.ld + ca;
    double w00, w01, w10 = 0.0, w11 = 0.0;
    if (INTERIOR && vec) {
      const double2 u = *reinterpret_cast<const double2*>(wrow);
      const double2 v = *reinterpret_cast<const double2*>(wrow + p.ld);
      w00 = u.x, w01 = u.y, w10 = v.x, w11 = v.y;
    } else {
      w00 = (oka && ca <= ra) ? wrow[0] : 0.0;
      w01 = (okb && cb <= ra) ? wrow[1] : 0.0;
      if (okr) {
        w10 = (oka && ca <= rb) ? wrow[p.ld] : 0.0;
        w11 = (okb && cb <= rb) ? wrow[p.ld + 1] : 0.0;
      }
    }
    const double b0 = sRow[rr * DP + D], b1 = sRow[(rr + 1) * DP + D];
    double f00 = 2.0, f01 = 2.0, f10 = 2.0, f11 = 2.0;
    if (!INTERIOR) {  // lower triangle only, diagonal once, nothing outside the matrix
      f00 = (!oka || ca > ra) ? 0.0 : (ca == ra ? 1.0 : 2.0);
      f01 = (!okb || cb > ra) ? 0.0 : (cb == ra ? 1.0 : 2.0);
      f10 = (!okr || !oka || ca > rb) ? 0.0 : (ca == rb ? 1.0 : 2.0);
      f11 = (!okr || !okb || cb > rb) ? 0.0 : (cb == rb ? 1.0 : 2.0);
    }
    w00 = (w00 - b0 * ba) * f00, w01 = (w01 - b0 * bb) * f01;
    w10 = (w10 - b1 * ba) * f10, w11 = (w11 - b1 * bb) * f11;
    double q00[D], q01[D], q10[D], q11[D];
    double s00 = 1e-300, s01 = 1e-300, s10 = 1e-300, s11 = 1e-300;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      if (DIM > 0 || i < dim) {
        const double x0 = sRow[rr * DP + i], x1 = sRow[(rr + 1) * DP + i];
        const double t00 = x0 - xa[i], t01 = x0 - xb[i], t10 = x1 - xa[i], t11 = x1 - xb[i];
        q00[i] = t00 * t00, q01[i] = t01 * t01, q10[i] = t10 * t10, q11[i] = t11 * t11;
        s00 += q00[i], s01 += q01[i], s10 += q10[i], s11 += q11[i];
      } else {
        q00[i] = q01[i] = q10[i] = q11[i] = 0.0;
      }
    }
    double v00, v01, v10, v11, e00, e01, e10, e11;
    radial_trace_factors<KIND>(s00, v00, e00);
    radial_trace_factors<KIND>(s01, v01, e01);
    radial_trace_factors<KIND>(s10, v10, e10);
    radial_trace_factors<KIND>(s11, v11, e11);
    acc[0] = fma(w00, v00, acc[0]);
    acc[0] = fma(w01, v01, acc[0]);
    acc[0] = fma(w10, v10, acc[0]);
    acc[0] = fma(w11, v11, acc[0]);
    e00 *= w00, e01 *= w01, e10 *= w10, e11 *= w11;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      acc[1 + i] = fma(e00, q00[i], acc[1 + i]);
      acc[1 + i] = fma(e01, q01[i], acc[1 + i]);
      acc[1 + i] = fma(e10, q10[i], acc[1 + i]);
      acc[1 + i] = fma(e11, q11[i], acc[1 + i]);
    }
  }
}

template <int KIND, int DIM>
__global__ void __launch_bounds__(FILL_THREADS, 2) kgrad_trace_kernel(const TraceParams p) {
  __shared__ double red[32];
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  __shared__ double sRowAll[8][8 * (D + 1)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int dim = p.dim;
  double inv[D], acc[D + 1];
#pragma unroll
  for (int i = 0; i < D; ++i) inv[i] = (DIM > 0 || i < dim) ? p.inv_len[i] : 0.0;
#pragma unroll
  for (int i = 0; i <= D; ++i) acc[i] = 0.0;

  // contiguous tile ranges per CTA, decoded once and advanced incrementally (row-major over the lower triangle)
  const long long per_cta = (p.ntiles + gridDim.x - 1) / gridDim.x;
  long long tile = blockIdx.x * per_cta;
  const long long tile_end = min(p.ntiles, tile + per_cta);
  long long ti = 0, tj = 0;
  if (tile < tile_end) tri_index(tile, ti, tj);  // tj <= ti
  for (; tile < tile_end; ++tile) {
    const bool interior = tj < ti && (ti + 1) * FT <= p.n;
    if (interior) trace_tile<KIND, DIM, true>(p, ti, tj, sRowAll[warp], inv, acc, lane, warp);
    else trace_tile<KIND, DIM, false>(p, ti, tj, sRowAll[warp], inv, acc, lane, warp);
    if (++tj > ti) tj = 0, ++ti;
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const double v = block_sum(acc[i], red);
    if (tid == 0 && (DIM > 0 || i <= dim)) p.partials[(long long)blockIdx.x * (dim + 1) + i] = v;
  }
}

// out[h] = scale[h] * sum_cta partials[cta][h]; fixed order -> deterministic.
__global__ void trace_reduce_kernel(const double* partials, int nctas, int H, const double* scale_dev, double* out) {
  __shared__ double red[32];
  for (int h = 0; h < H; ++h) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nctas; i += blockDim.x) s += partials[(long long)i * H + h];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[h] = scale_dev != nullptr ? s * scale_dev[h] : s;
    __syncthreads();
  }
}

// Rectangular block of the same trace (multi-GPU block-cyclic layout, fvgp_b200/sharded.py): W is an m x n block of
// KV^-1 whose rows belong to points x1 / entries b1 and columns to x2 / b2.  The first `diag_rows` rows form a
// diagonal block aligned with the columns (only its lower triangle counts, the diagonal once); every other entry
// stands for itself and its mirror image (weight 2).
struct TraceBlockParams {
  const double* x1;
  const double* x2;
  const double* W;
  const double* b1;
  const double* b2;
  double* partials;
  long long m, n, ld, tiles_j, ntiles;
  long long diag_rows;
  double inv_len[kMaxDim];
  int dim;
};

template <int DIM>
__global__ void __launch_bounds__(FILL_THREADS) kgrad_trace_block_kernel(const TraceBlockParams p) {
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  const int dim = p.dim;
  const double sqrt3 = 1.7320508075688772;
  double inv[D], acc[D + 1];
#pragma unroll
  for (int i = 0; i < D; ++i) inv[i] = (DIM > 0 || i < dim) ? p.inv_len[i] : 0.0;
#pragma unroll
  for (int i = 0; i <= D; ++i) acc[i] = 0.0;
  for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long ti = tile / p.tiles_j, tj = tile - ti * p.tiles_j;
    const long long r0 = ti * FT + warp * 8, c0 = tj * FT;
    if (r0 + 7 < p.diag_rows && c0 > r0 + 7) continue;  // strictly above the diagonal of the diagonal block
    const long long cc[2] = {c0 + lane, c0 + lane + 32};
    double xc[2][D], bc[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const bool ok = cc[q] < p.n;
      bc[q] = ok ? p.b2[cc[q]] : 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) xc[q][i] = (ok && (DIM > 0 || i < dim)) ? p.x2[cc[q] * dim + i] : 0.0;
    }
    for (int rr = 0; rr < 8; ++rr) {
      const long long r = r0 + rr;
      if (r >= p.m) break;
      const double br = p.b1[r];
      double xr[D];
#pragma unroll
      for (int i = 0; i < D; ++i) xr[i] = (DIM > 0 || i < dim) ? p.x1[r * dim + i] : 0.0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const long long c = cc[q];
        if (c >= p.n) continue;
        double weight = 2.0;
        if (r < p.diag_rows) {
          if (c > r) continue;
          if (c == r) weight = 1.0;
        }
        const double w = (p.W[r * p.ld + c] - br * bc[q]) * weight;
        double t2[D], s = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const double t = (xr[i] - xc[q][i]) * inv[i];
          t2[i] = t * t;
          s += t2[i];
        }
        const double a = sqrt3 * sqrt_pos(fmax(s, 1e-300));
        const double ea = exp_neg(fmin(a, 1.0e9));
        const double wea = w * ea;
        acc[0] = fma(wea, 1.0 + a, acc[0]);
#pragma unroll
        for (int i = 0; i < D; ++i) acc[1 + i] = fma(wea, t2[i], acc[1 + i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const double v = block_sum(acc[i], red);
    if (tid == 0 && (DIM > 0 || i <= dim)) p.partials[(long long)blockIdx.x * (dim + 1) + i] = v;
  }
}

// The same block trace for every radial family of the fused fill (squared exponential, exponential, Matern-3/2,
// Matern-5/2): coordinates arrive scaled by inv_scale_i * c_KIND / length (see radial_trace_factors), the raw sums
//   acc[0] = sum W f(u),   acc[1+i] = sum W h(u) q_i
// are converted to the traces against (amp, inv_scale_1..D, length) on the host after the all-reduce.
template <int KIND, int DIM>
__global__ void __launch_bounds__(FILL_THREADS) kgrad_trace_block_radial_kernel(const TraceBlockParams p) {
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int D = DIM > 0 ? DIM : kMaxDim;
  const int dim = p.dim;
  double inv[D], acc[D + 1];
#pragma unroll
  for (int i = 0; i < D; ++i) inv[i] = (DIM > 0 || i < dim) ? p.inv_len[i] : 0.0;
#pragma unroll
  for (int i = 0; i <= D; ++i) acc[i] = 0.0;
  for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const long long ti = tile / p.tiles_j, tj = tile - ti * p.tiles_j;
    const long long r0 = ti * FT + warp * 8, c0 = tj * FT;
    if (r0 + 7 < p.diag_rows && c0 > r0 + 7) continue;  // strictly above the diagonal of the diagonal block
    const long long cc[2] = {c0 + lane, c0 + lane + 32};
    double xc[2][D], bc[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const bool ok = cc[q] < p.n;
      bc[q] = ok ? p.b2[cc[q]] : 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) xc[q][i] = (ok && (DIM > 0 || i < dim)) ? p.x2[cc[q] * dim + i] * inv[i] : 0.0;
    }
    for (int rr = 0; rr < 8; ++rr) {
      const long long r = r0 + rr;
      if (r >= p.m) break;
      const double br = p.b1[r];
      double xr[D];
#pragma unroll
      for (int i = 0; i < D; ++i) xr[i] = (DIM > 0 || i < dim) ? p.x1[r * dim + i] * inv[i] : 0.0;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const long long c = cc[q];
        if (c >= p.n) continue;
        double weight = 2.0;
        if (r < p.diag_rows) {
          if (c > r) continue;
          if (c == r) weight = 1.0;
        }
        const double w = (p.W[r * p.ld + c] - br * bc[q]) * weight;
        double t2[D], s = 1e-300;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          const double t = xr[i] - xc[q][i];
          t2[i] = t * t;
          s += t2[i];
        }
        double f, h;
        radial_trace_factors<KIND>(s, f, h);
        acc[0] = fma(w, f, acc[0]);
        const double wh = w * h;
#pragma unroll
        for (int i = 0; i < D; ++i) acc[1 + i] = fma(wh, t2[i], acc[1 + i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i <= D; ++i) {
    const double v = block_sum(acc[i], red);
    if (tid == 0 && (DIM > 0 || i <= dim)) p.partials[(long long)blockIdx.x * (dim + 1) + i] = v;
  }
}

// accum[h] += sum_cta partials[cta][h]   (raw sums; the host applies the chain-rule factors)
__global__ void trace_accumulate_raw_kernel(const double* partials, int nctas, int H, double* accum) {
  __shared__ double red[32];
  for (int h = 0; h < H; ++h) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nctas; i += blockDim.x) s += partials[(long long)i * H + h];
    s = block_sum(s, red);
    if (threadIdx.x == 0) accum[h] += s;
    __syncthreads();
  }
}

// accum[h] += scale[h] * sum_cta partials[cta][h]
__global__ void trace_accumulate_kernel(const double* partials, int nctas, int H, double amp, const double* inv_len_dev,
                                        double* accum) {
  __shared__ double red[32];
  for (int h = 0; h < H; ++h) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nctas; i += blockDim.x) s += partials[(long long)i * H + h];
    s = block_sum(s, red);
    if (threadIdx.x == 0) accum[h] += (h == 0) ? s : s * 3.0 * amp * inv_len_dev[h - 1];
    __syncthreads();
  }
}

// Radial kernel applied elementwise to a user-supplied distance array (the fvgp.kernels
// functions called on a plain ndarray instead of on get_distance_matrix's result).
template <int KIND>
__global__ void radial_elementwise_kernel(const double* __restrict__ d, long long count, double amp, double c_arg,
                                          double c_aux, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const double v = d[i];
    out[i] = radial_value<KIND, false>(fmax(v * v, 1e-300), amp, c_arg, c_aux);
  }
}

// sum_ij (Kinv - b b^T)_ij * dK_ij for a materialised symmetric dK (user kernels with their own
// gradient function); only the lower triangles are read.  One partial per CTA.
__global__ void __launch_bounds__(256) trace_sym_product_kernel(const double* __restrict__ Kinv, long long ld,
                                                                const double* __restrict__ b,
                                                                const double* __restrict__ dK, long long lddk,
                                                                long long n, double* partials) {
  __shared__ double red[32];
  double acc = 0.0;
  for (long long r = blockIdx.x; r < n; r += gridDim.x) {
    const double br = b[r];
    for (long long c = threadIdx.x; c <= r; c += blockDim.x) {
      const double w = (Kinv[r * ld + c] - br * b[c]) * ((c == r) ? 1.0 : 2.0);
      acc = fma(w, dK[r * lddk + c], acc);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void sum_partials_kernel(const double* partials, int count, double* out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) s += partials[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

__global__ void kgrad_dense_kernel(const double* __restrict__ x1, long long n1, const double* __restrict__ x2,
                                   long long n2, int dim, double amp, const double* __restrict__ len, double* out) {
  const long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= n2) return;
  const double sqrt3 = 1.7320508075688772;
  const long long plane = n1 * n2;
  for (long long r = blockIdx.y; r < n1; r += gridDim.y) {  // grid.y is capped at 65535: rows are strided
    double s = 0.0, dx2[kMaxDim];
    for (int i = 0; i < dim; ++i) {
      const double dx = fabs(x1[r * dim + i] - x2[c * dim + i]);
      const double t = dx / len[i];
      dx2[i] = dx * dx;
      s += t * t;
    }
    const double d = sqrt(s), a = sqrt3 * d, ea = exp(-a);
    const long long at = r * n2 + c;
    out[at] = (1.0 + a) * ea;
    for (int i = 0; i < dim; ++i) {
      double v = 0.0;
      if (d != 0.0) {
        const double dadl = sqrt3 * (-dx2[i] / (len[i] * len[i] * len[i] * d));
        v = amp * (dadl * ea - (1.0 + a) * dadl * ea);  // gp_prior.py:431-434, kernels.py:137-141
      }
      out[(1 + i) * plane + at] = v;
    }
  }
}

}  // namespace fvgp

using namespace fvgp;

// Kind-specific constants of the fill (shared by the single and the batched entry points).
static bool fill_constants(int kind, double amp, double length, bool centred, double& c_arg, double& c_aux, double& fold) {
  c_arg = 0.0, c_aux = 0.0, fold = 1.0;
  switch (kind) {
    case FVGP_K_MATERN32: c_arg = sqrt(3.0) / length; fold = c_arg; break;
    case FVGP_K_MATERN52:
      c_arg = sqrt(5.0) / length, fold = c_arg;
      c_aux = centred ? amp / 3.0 : amp * 5.0 / (3.0 * length * length);
      break;
    case FVGP_K_MATERN52_ROBUST:  // same evaluation, quadratic coefficient 15 / length^2 (= 3 a^2 in scaled coordinates)
      c_arg = sqrt(5.0) / length, fold = c_arg;
      c_aux = centred ? amp * 3.0 : amp * 15.0 / (length * length);
      break;
    case FVGP_K_SQEXP: c_arg = 1.0 / (2.0 * length * length); fold = sqrt(c_arg); break;
    case FVGP_K_EXP: c_arg = 1.0 / length; fold = c_arg; break;
    case FVGP_K_WENDLAND: c_arg = 1.0 / length; fold = c_arg; break;
    case FVGP_K_DISTANCE: break;
    default: return false;
  }
  return true;
}

// Lower-triangle fills of `batch` proposals in ONE launch (population evaluator).  d_thetas: device scratch of
// kfill_batch_theta_len(batch) doubles; matrix b at d_K + b * kstride.  Enqueue only.
int64_t kfill_batch_theta_len(int batch) { return (int64_t)batch * (int64_t)(sizeof(FillTheta) / sizeof(double)); }

int kfill_lower_batch_enqueue(int kind, const double* d_x, int64_t n, int dim, int batch, const double* h_amp,
                              const double* h_inv_scale, const double* h_length, const double* h_centre,
                              const double* d_noise, double* d_K, int64_t ldk, int64_t kstride, double* d_thetas,
                              cudaStream_t st) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim && n > 0 && ldk >= n && batch >= 1 && batch <= 65535);
  FillParams p;
  p.x1 = d_x, p.x2 = d_x, p.noise = d_noise, p.K = d_K;
  p.n1 = n, p.n2 = n, p.ldk = ldk, p.amp = 0.0, p.dim = dim, p.mode = FVGP_FILL_LOWER;
  p.tiles_i = (n + FT - 1) / FT;
  p.tiles_j = p.tiles_i;
  p.ntiles = p.tiles_i * (p.tiles_i + 1) / 2;
  p.vec2 = (ldk % 2 == 0 && kstride % 2 == 0 && ((uintptr_t)d_K % 16 == 0)) ? 1 : 0;
  p.bulk = 0, p.band = 0, p.c_arg = 0.0, p.c_aux = 0.0;
  const bool centred = h_centre != nullptr && dim <= 4;
  std::vector<FillTheta> th((size_t)batch);
  for (int b = 0; b < batch; ++b) {
    double fold = 1.0;
    th[b].amp = h_amp[b];
    FVGP_REQUIRE(fill_constants(kind, h_amp[b], h_length[b], centred, th[b].c_arg, th[b].c_aux, fold));
    for (int i = 0; i < kMaxDim; ++i)
      th[b].inv_scale[i] = i < dim ? h_inv_scale[(size_t)b * dim + i] * (centred ? fold : 1.0) : 0.0;
  }
  for (int i = 0; i < kMaxDim; ++i) {
    p.inv_scale[i] = 0.0;
    p.centre[i] = (centred && i < dim) ? h_centre[i] : 0.0;
  }
  // pageable source: the runtime stages the bytes before returning, so the vector may go out of scope
  FVGP_CUDA_OK(cudaMemcpyAsync(d_thetas, th.data(), sizeof(FillTheta) * batch, cudaMemcpyHostToDevice, st));
  const FillTheta* dth = reinterpret_cast<const FillTheta*>(d_thetas);
  switch (kind) {
    case FVGP_K_MATERN32: return launch_fill_batch_dim<FVGP_K_MATERN32>(p, centred, dth, kstride, batch, st);
    case FVGP_K_MATERN52:
    case FVGP_K_MATERN52_ROBUST: return launch_fill_batch_dim<FVGP_K_MATERN52>(p, centred, dth, kstride, batch, st);
    case FVGP_K_SQEXP: return launch_fill_batch_dim<FVGP_K_SQEXP>(p, centred, dth, kstride, batch, st);
    case FVGP_K_EXP: return launch_fill_batch_dim<FVGP_K_EXP>(p, centred, dth, kstride, batch, st);
    case FVGP_K_WENDLAND: return launch_fill_batch_dim<FVGP_K_WENDLAND>(p, centred, dth, kstride, batch, st);
    default: return launch_fill_batch_dim<FVGP_K_DISTANCE>(p, centred, dth, kstride, batch, st);
  }
}

extern "C" {

// test / profiling hook: mirror tile of the symmetric fill through 0 = plain coalesced stores (CTA barrier),
// 1 = one 512-byte TMA bulk store per row (CTA barrier, double-buffered staging; default), 2 = per-warp 64-byte bulk
// stores without any CTA barrier (A/B: slower, the 64-byte pieces cost more in the store path than the barriers did)
int fvgp_set_bulk_store(int on) {
  const int old = g_use_bulk_store;
  g_use_bulk_store = on < 0 ? 0 : (on > 2 ? 2 : on);
  return old;
}

int fvgp_kfill_dense(int kind, int mode, const double* d_x1, int64_t n1, const double* d_x2, int64_t n2, int dim,
                     double amp, const double* h_inv_scale, const double* h_centre, double length,
                     const double* d_noise, double* d_K, int64_t ldk, void* stream) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim && n1 >= 0 && n2 >= 0 && ldk >= n2);
  FVGP_REQUIRE(mode == FVGP_FILL_FULL || (n1 == n2));
  if (n1 == 0 || n2 == 0) return 0;
  FillParams p;
  p.x1 = d_x1, p.x2 = d_x2, p.noise = d_noise, p.K = d_K;
  p.n1 = n1, p.n2 = n2, p.ldk = ldk, p.amp = amp, p.dim = dim, p.mode = mode;
  p.tiles_i = (n1 + FT - 1) / FT;
  p.tiles_j = (n2 + FT - 1) / FT;
  p.ntiles = mode == FVGP_FILL_FULL ? p.tiles_i * p.tiles_j : p.tiles_i * (p.tiles_i + 1) / 2;
  p.vec2 = (ldk % 2 == 0 && ((uintptr_t)d_K % 16 == 0)) ? 1 : 0;
  p.bulk = (g_use_bulk_store && mode == FVGP_FILL_SYMMETRIC && p.vec2) ? g_use_bulk_store : 0;
  {
    static int band = -1;  // FVGP_FILL_BAND: tile rows per band of the symmetric fill's tile order (0 = column sweep)
    if (band < 0) {
      const char* e = getenv("FVGP_FILL_BAND");
      band = e ? atoi(e) : 0;  // measured on B200: the column sweep is as fast or faster (4.6 vs 4.4 TB/s at N = 50k)
      if (band < 0 || band > 64) band = 0;
    }
    p.band = band;
  }
  p.c_arg = 0.0, p.c_aux = 0.0;
  // Centred ("whitened") fast path: coordinates become (x - centre) * inv_scale * c with the kind's argument
  // constant c folded in.  The caller vouches (by passing h_centre) that |x - centre| * inv_scale * c <= 512
  // for every point, which bounds the extra relative error of an entry by ~2e-13.
  const bool centred = h_centre != nullptr && dim <= 4;
  double fold = 1.0;
  FVGP_REQUIRE(fill_constants(kind, amp, length, centred, p.c_arg, p.c_aux, fold));
  for (int i = 0; i < kMaxDim; ++i) {
    p.inv_scale[i] = i < dim ? h_inv_scale[i] * (centred ? fold : 1.0) : 0.0;
    p.centre[i] = (centred && i < dim) ? h_centre[i] : 0.0;
  }
  cudaStream_t st = (cudaStream_t)stream;
  switch (kind) {
    case FVGP_K_MATERN32: return launch_fill_dim<FVGP_K_MATERN32>(p, centred, st);
    case FVGP_K_MATERN52:
    case FVGP_K_MATERN52_ROBUST: return launch_fill_dim<FVGP_K_MATERN52>(p, centred, st);
    case FVGP_K_SQEXP: return launch_fill_dim<FVGP_K_SQEXP>(p, centred, st);
    case FVGP_K_EXP: return launch_fill_dim<FVGP_K_EXP>(p, centred, st);
    case FVGP_K_WENDLAND: return launch_fill_dim<FVGP_K_WENDLAND>(p, centred, st);
    default: return launch_fill_dim<FVGP_K_DISTANCE>(p, centred, st);
  }
}

int fvgp_radial_elementwise(int kind, const double* d_dist, int64_t count, double amp, double length, double* d_out,
                            void* stream) {
  if (count <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const long long want = (count + 255) / 256, cap = (long long)sm_count() * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  switch (kind) {
    case FVGP_K_MATERN32:
      launch(radial_elementwise_kernel<FVGP_K_MATERN32>, grid, 256, 0, st, d_dist, count, amp, sqrt(3.0) / length, 0.0, d_out);
      break;
    case FVGP_K_MATERN52:
      launch(radial_elementwise_kernel<FVGP_K_MATERN52>, grid, 256, 0, st, d_dist, count, amp, sqrt(5.0) / length,
                                                                        amp * 5.0 / (3.0 * length * length), d_out);
      break;
    case FVGP_K_MATERN52_ROBUST:
      launch(radial_elementwise_kernel<FVGP_K_MATERN52>, grid, 256, 0, st, d_dist, count, amp, sqrt(5.0) / length,
                                                                        amp * 15.0 / (length * length), d_out);
      break;
    case FVGP_K_SQEXP:
      launch(radial_elementwise_kernel<FVGP_K_SQEXP>, grid, 256, 0, st, d_dist, count, amp, 1.0 / (2.0 * length * length), 0.0, d_out);
      break;
    case FVGP_K_EXP:
      launch(radial_elementwise_kernel<FVGP_K_EXP>, grid, 256, 0, st, d_dist, count, amp, 1.0 / length, 0.0, d_out);
      break;
    case FVGP_K_WENDLAND:
      launch(radial_elementwise_kernel<FVGP_K_WENDLAND>, grid, 256, 0, st, d_dist, count, amp, 1.0 / length, 0.0, d_out);
      break;
    default: FVGP_REQUIRE(!"unknown kernel kind");
  }
  FVGP_LAUNCH_OK();
  return 0;
}

static inline long long trace_grid(int64_t n) {
  const long long t = (n + FT - 1) / FT, ntiles = t * (t + 1) / 2;
  const long long cap = (long long)sm_count() * 2;  // 126 registers: two resident CTAs per SM, one wave
  return ntiles < cap ? ntiles : cap;
}

int64_t fvgp_kgrad_partials_len(int64_t n, int dim) { return (trace_grid(n) + 2) * (dim + 1) + kMaxDim; }

}  // extern "C"

template <int KIND>
static void launch_trace_kind(const TraceParams& p, unsigned grid, cudaStream_t st) {
  switch (p.dim) {
    case 1: launch(kgrad_trace_kernel<KIND, 1>, grid, FILL_THREADS, 0, st, p); break;
    case 2: launch(kgrad_trace_kernel<KIND, 2>, grid, FILL_THREADS, 0, st, p); break;
    case 3: launch(kgrad_trace_kernel<KIND, 3>, grid, FILL_THREADS, 0, st, p); break;
    case 4: launch(kgrad_trace_kernel<KIND, 4>, grid, FILL_THREADS, 0, st, p); break;
    default: launch(kgrad_trace_kernel<KIND, 0>, grid, FILL_THREADS, 0, st, p); break;
  }
}

// Shared driver: coordinates scaled by h_coord_scale[i]; out[h] = h_out_scale[h] * (raw sum h).  Enqueue only:
// the H results land in d_out (device; nullptr = the tail of d_partials) with no host synchronisation, so that
// the population evaluator (dense_linalg.cu) can run many of these on concurrent streams.
int trace_radial_enqueue(int kind, const double* d_x, int64_t n, int dim, const double* h_coord_scale,
                         const double* h_out_scale, const double* d_Kinv, int64_t ld, const double* d_b,
                         double* d_partials, double* d_out, cudaStream_t st) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim && n > 0);
  TraceParams p;
  p.x = d_x, p.Kinv = d_Kinv, p.b = d_b, p.partials = d_partials, p.n = n, p.ld = ld, p.dim = dim;
  const long long t = (n + FT - 1) / FT;
  p.ntiles = t * (t + 1) / 2;
  for (int i = 0; i < kMaxDim; ++i) p.inv_len[i] = i < dim ? h_coord_scale[i] : 0.0;
  const unsigned grid = (unsigned)trace_grid(n);
  switch (kind) {
    case FVGP_K_MATERN32: launch_trace_kind<FVGP_K_MATERN32>(p, grid, st); break;
    case FVGP_K_MATERN52: launch_trace_kind<FVGP_K_MATERN52>(p, grid, st); break;
    case FVGP_K_SQEXP: launch_trace_kind<FVGP_K_SQEXP>(p, grid, st); break;
    case FVGP_K_EXP: launch_trace_kind<FVGP_K_EXP>(p, grid, st); break;
    default: FVGP_REQUIRE(!"gradient traces exist for the Matern-3/2, Matern-5/2, squared-exponential and exponential kinds");
  }
  FVGP_LAUNCH_OK();
  const int H = dim + 1;
  double* d_tail = d_partials + (long long)grid * H;
  double* d_scale = d_tail + H;
  if (d_out == nullptr) d_out = d_tail;
  if (h_out_scale != nullptr) {
    FVGP_CUDA_OK(cudaMemcpyAsync(d_scale, h_out_scale, H * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
    d_scale = nullptr;  // raw sums
  }
  launch(trace_reduce_kernel, 1, 256, 0, st, d_partials, (int)grid, H, d_scale, d_out);
  FVGP_LAUNCH_OK();
  return 0;
}

static int trace_radial_impl(int kind, const double* d_x, int64_t n, int dim, const double* h_coord_scale,
                             const double* h_out_scale, const double* d_Kinv, int64_t ld, const double* d_b,
                             double* d_partials, double* h_out, cudaStream_t st) {
  int r = trace_radial_enqueue(kind, d_x, n, dim, h_coord_scale, h_out_scale, d_Kinv, ld, d_b, d_partials, nullptr, st);
  if (r != 0) return r;
  const int H = dim + 1;
  const double* d_out = d_partials + (long long)trace_grid(n) * H;
  FVGP_CUDA_OK(cudaMemcpyAsync(h_out, d_out, H * sizeof(double), cudaMemcpyDeviceToHost, st));
  FVGP_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

extern "C" {

int fvgp_kgrad_trace_matern32(const double* d_x, int64_t n, int dim, const double* h_theta, const double* d_Kinv,
                              int64_t ld, const double* d_b, double* d_partials, double* h_out, void* stream) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim);
  // default kernel, theta = (amp, l_1..l_dim): dK/dl_i = amp * 3 t_i^2 e^-a / l_i with t_i = dx_i / l_i; the kernel
  // accumulates e^-a * (sqrt(3) t_i)^2, hence the factor amp / l_i
  double coord[kMaxDim], scale[kMaxDim + 1];
  scale[0] = 1.0;
  for (int i = 0; i < dim; ++i) {
    coord[i] = sqrt(3.0) / h_theta[1 + i];
    scale[1 + i] = h_theta[0] / h_theta[1 + i];
  }
  return trace_radial_impl(FVGP_K_MATERN32, d_x, n, dim, coord, scale, d_Kinv, ld, d_b, d_partials, h_out,
                           (cudaStream_t)stream);
}

int fvgp_kgrad_trace_radial(int kind, const double* d_x, int64_t n, int dim, double amp, const double* h_inv_scale,
                            double length, const double* d_Kinv, int64_t ld, const double* d_b, double* d_partials,
                            double* h_out, void* stream) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim && length > 0.0);
  double fold = 1.0;
  switch (kind) {
    case FVGP_K_MATERN32: fold = sqrt(3.0) / length; break;
    case FVGP_K_MATERN52: fold = sqrt(5.0) / length; break;
    case FVGP_K_SQEXP: fold = sqrt(0.5) / length; break;
    case FVGP_K_EXP: fold = 1.0 / length; break;
    default: FVGP_REQUIRE(!"gradient traces exist for the Matern-3/2, Matern-5/2, squared-exponential and exponential kinds");
  }
  // raw sums R_0 = sum W f, R_i = sum W h q_i  ->  h_out = (T_amp, T_s1..T_sD, T_length):
  //   T_amp = R_0,  T_si = -(amp / s_i) R_i,  T_length = (amp / length) sum_i R_i
  double coord[kMaxDim], scale[kMaxDim + 1], raw[kMaxDim + 1];
  scale[0] = 1.0;
  for (int i = 0; i < dim; ++i) coord[i] = h_inv_scale[i] * fold, scale[1 + i] = 1.0;
  int r = trace_radial_impl(kind, d_x, n, dim, coord, scale, d_Kinv, ld, d_b, d_partials, raw, (cudaStream_t)stream);
  if (r != 0) return r;
  double sum = 0.0;
  h_out[0] = raw[0];
  for (int i = 0; i < dim; ++i) {
    sum += raw[1 + i];
    h_out[1 + i] = h_inv_scale[i] != 0.0 ? -(amp / h_inv_scale[i]) * raw[1 + i] : 0.0;
  }
  h_out[dim + 1] = (amp / length) * sum;
  return 0;
}

int64_t fvgp_kgrad_block_partials_len(int dim) { return (int64_t)(sm_count() * 4 + 2) * (dim + 1) + kMaxDim; }

int fvgp_kgrad_trace_block_matern32(const double* d_x1, int64_t m, const double* d_x2, int64_t n, int dim,
                                    const double* h_theta, const double* d_W, int64_t ldw, const double* d_b1,
                                    const double* d_b2, int64_t diag_rows, double* d_partials, double* d_accum,
                                    void* stream) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim && m >= 0 && n >= 0 && diag_rows >= 0 && diag_rows <= m);
  if (m == 0 || n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  TraceBlockParams p;
  p.x1 = d_x1, p.x2 = d_x2, p.W = d_W, p.b1 = d_b1, p.b2 = d_b2, p.partials = d_partials;
  p.m = m, p.n = n, p.ld = ldw, p.dim = dim, p.diag_rows = diag_rows;
  p.tiles_j = (n + FT - 1) / FT;
  p.ntiles = ((m + FT - 1) / FT) * p.tiles_j;
  double inv_len[kMaxDim];
  for (int i = 0; i < kMaxDim; ++i) inv_len[i] = p.inv_len[i] = i < dim ? 1.0 / h_theta[1 + i] : 0.0;
  const long long cap = (long long)sm_count() * 4;
  const unsigned grid = (unsigned)(p.ntiles < cap ? p.ntiles : cap);
  switch (dim) {
    case 1: launch(kgrad_trace_block_kernel<1>, grid, FILL_THREADS, 0, st, p); break;
    case 2: launch(kgrad_trace_block_kernel<2>, grid, FILL_THREADS, 0, st, p); break;
    case 3: launch(kgrad_trace_block_kernel<3>, grid, FILL_THREADS, 0, st, p); break;
    case 4: launch(kgrad_trace_block_kernel<4>, grid, FILL_THREADS, 0, st, p); break;
    default: launch(kgrad_trace_block_kernel<0>, grid, FILL_THREADS, 0, st, p); break;
  }
  FVGP_LAUNCH_OK();
  const int H = dim + 1;
  double* d_invlen = d_partials + (long long)(sm_count() * 4 + 2) * H;
  FVGP_CUDA_OK(cudaMemcpyAsync(d_invlen, inv_len, dim * sizeof(double), cudaMemcpyHostToDevice, st));
  launch(trace_accumulate_kernel, 1, 256, 0, st, d_partials, (int)grid, H, h_theta[0], d_invlen, d_accum);
  FVGP_LAUNCH_OK();
  return 0;
}

}  // extern "C" (templates need C++ linkage)

template <int KIND>
static void launch_trace_block_radial(const TraceBlockParams& p, unsigned grid, cudaStream_t st) {
  switch (p.dim) {
    case 1: launch(kgrad_trace_block_radial_kernel<KIND, 1>, grid, FILL_THREADS, 0, st, p); break;
    case 2: launch(kgrad_trace_block_radial_kernel<KIND, 2>, grid, FILL_THREADS, 0, st, p); break;
    case 3: launch(kgrad_trace_block_radial_kernel<KIND, 3>, grid, FILL_THREADS, 0, st, p); break;
    case 4: launch(kgrad_trace_block_radial_kernel<KIND, 4>, grid, FILL_THREADS, 0, st, p); break;
    default: launch(kgrad_trace_block_radial_kernel<KIND, 0>, grid, FILL_THREADS, 0, st, p); break;
  }
}

extern "C" {

int fvgp_kgrad_trace_block_radial(int kind, const double* d_x1, int64_t m, const double* d_x2, int64_t n, int dim,
                                  const double* h_inv_scale, double length, const double* d_W, int64_t ldw,
                                  const double* d_b1, const double* d_b2, int64_t diag_rows, double* d_partials,
                                  double* d_accum_raw, void* stream) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim && m >= 0 && n >= 0 && diag_rows >= 0 && diag_rows <= m && length > 0.0);
  double fold = 1.0;
  switch (kind) {
    case FVGP_K_MATERN32: fold = sqrt(3.0) / length; break;
    case FVGP_K_MATERN52: fold = sqrt(5.0) / length; break;
    case FVGP_K_SQEXP: fold = sqrt(0.5) / length; break;
    case FVGP_K_EXP: fold = 1.0 / length; break;
    default: FVGP_REQUIRE(!"gradient traces exist for the Matern-3/2, Matern-5/2, squared-exponential and exponential kinds");
  }
  if (m == 0 || n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  TraceBlockParams p;
  p.x1 = d_x1, p.x2 = d_x2, p.W = d_W, p.b1 = d_b1, p.b2 = d_b2, p.partials = d_partials;
  p.m = m, p.n = n, p.ld = ldw, p.dim = dim, p.diag_rows = diag_rows;
  p.tiles_j = (n + FT - 1) / FT;
  p.ntiles = ((m + FT - 1) / FT) * p.tiles_j;
  for (int i = 0; i < kMaxDim; ++i) p.inv_len[i] = i < dim ? h_inv_scale[i] * fold : 0.0;
  const long long cap = (long long)sm_count() * 4;
  const unsigned grid = (unsigned)(p.ntiles < cap ? p.ntiles : cap);
  switch (kind) {
    case FVGP_K_MATERN32: launch_trace_block_radial<FVGP_K_MATERN32>(p, grid, st); break;
    case FVGP_K_MATERN52: launch_trace_block_radial<FVGP_K_MATERN52>(p, grid, st); break;
    case FVGP_K_SQEXP: launch_trace_block_radial<FVGP_K_SQEXP>(p, grid, st); break;
    default: launch_trace_block_radial<FVGP_K_EXP>(p, grid, st); break;
  }
  FVGP_LAUNCH_OK();
  launch(trace_accumulate_raw_kernel, 1, 256, 0, st, (const double*)d_partials, (int)grid, dim + 1, d_accum_raw);
  FVGP_LAUNCH_OK();
  return 0;
}

int fvgp_trace_sym_product(const double* d_Kinv, int64_t ld, const double* d_b, const double* d_dK, int64_t lddk,
                           int64_t n, double* d_partials, double* h_out, void* stream) {
  FVGP_REQUIRE(n > 0);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)(n < (int64_t)sm_count() * 8 ? n : (int64_t)sm_count() * 8);
  launch(trace_sym_product_kernel, grid, 256, 0, st, d_Kinv, ld, d_b, d_dK, lddk, n, d_partials);
  launch(sum_partials_kernel, 1, 256, 0, st, d_partials, grid, d_partials + grid);
  FVGP_LAUNCH_OK();
  FVGP_CUDA_OK(cudaMemcpyAsync(h_out, d_partials + grid, sizeof(double), cudaMemcpyDeviceToHost, st));
  FVGP_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int fvgp_kgrad_dense_matern32(const double* d_x1, int64_t n1, const double* d_x2, int64_t n2, int dim,
                              const double* h_theta, double* d_out, void* stream) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim && n1 > 0 && n2 > 0 && (n2 + 255) / 256 <= 0x7fffffffll);
  cudaStream_t st = (cudaStream_t)stream;
  double* d_len = nullptr;  // the length scales travel through a small stream-ordered device buffer
  FVGP_CUDA_OK(cudaMallocAsync((void**)&d_len, kMaxDim * sizeof(double), st));
  cudaError_t err = cudaMemcpyAsync(d_len, h_theta + 1, dim * sizeof(double), cudaMemcpyHostToDevice, st);
  if (err == cudaSuccess) {
    dim3 grid((unsigned)((n2 + 255) / 256), (unsigned)(n1 < 65535 ? n1 : 65535));
    launch(kgrad_dense_kernel, grid, 256, 0, st, d_x1, n1, d_x2, n2, dim, h_theta[0], d_len, d_out);
    err = cudaGetLastError();
  }
  cudaFreeAsync(d_len, st);  // released on every path
  FVGP_CUDA_OK(err);
  return 0;
}

}  // extern "C"
