// gp2Scale on the device: compact-support (Wendland) covariance straight to canonical CSR,
// CSR SpMV, preconditioned CG and the Lanczos recurrences of the SLQ log-determinant.
//
// Reference behaviour being replaced (lbl-camera/fvGP):
//   kernels.py:502-528          wendland_anisotropic_gp2Scale_cpu   (dense B x B block per dask task)
//   gp2Scale_covariance.py:136-170, :240-287   np.nonzero pattern, triu mask, mirror, COO -> CSR sort
//   gp_kv.py:655-661            K + diag(V) via setdiag
//   gp_lin_alg.py:1213-1291     calculate_sparse_conj_grad -> scipy.sparse.linalg.cg
//   gp_lin_alg.py:604-622       block-Jacobi preconditioner
//   gp_lin_alg.py:1103-1181     calculate_random_logdet -> imate SLQ (device part: Lanczos)
//
// Bit-exact sparsity: an entry is stored iff the reference's value is non-zero, which is
// s < 1 for s = sum_i ((x1_i - x2_i) / theta_i)^2 accumulated in axis order with every
// operation rounded separately (numpy never fuses).  The predicate below uses the
// __dsub_rn/__ddiv_rn/__dmul_rn/__dadd_rn intrinsics, which ptxas may not contract.
// Bounding-box culls (tile, super tile) and the per-pair test evaluate an approximate s (reciprocal
// multiply + FMA): >= 1 + 1e-12 is certainly outside, < 1 - 1e-12 certainly inside (the approximation
// error is a dozen ulp); only the sliver in between runs the exact sequence, so neither a stored entry
// can be lost nor a spurious one added.  Values always come from the exact sequence (separate kernel).
//
// Layout: 32-point row tiles (one warp each) against 32-point column tiles, column tiles
// grouped into super tiles of 32 for a two-level box cull.  Inside a surviving tile pair
// the warp walks the rows; lanes hold the 32 columns, hits are compacted with
// ballot + popc prefix so each row's entries land in ascending column order -- the CSR
// is canonical by construction, no sort, no atomics, deterministic.  Pipeline: count pass ->
// exclusive scan -> index pass (same geometry, writes column indices) -> value pass (lane-dense).
#include "../../include/fvgp_b200.h"
#include "common.cuh"
#include "comm.cuh"
#include <algorithm>
#include <cstdlib>

namespace fvgp {

constexpr int WT = 32;          // points per tile
constexpr int WS = 32;          // tiles per super tile
constexpr int W_WARPS = 4;      // row tiles per CTA

struct WendlandParams {
  const double* x1;
  const double* x2;
  const double* aabb1;
  const double* aabb2;
  const long long* indptr;
  const double* noise;
  long long* rowcount;
  long long* stats;
  int* chunk;
  long long supers_per_chunk;
  int n_chunks;
  int* indices;
  double* data;
  long long n1, n2, tiles1, tiles2, super2;
  long long row0;  // global index of row 0 of x1 inside x2's numbering (row slabs of a symmetric matrix); diagonal = row + row0
  double amp;
  double theta[kMaxDim];
  int dim;
};

// Bounding boxes: tile t -> lo[dim], hi[dim] at aabb + t*2*dim; super tiles follow the tiles.
__global__ void aabb_tiles_kernel(const double* __restrict__ x, long long n, int dim, double* aabb, long long tiles) {
  const long long t = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= tiles) return;
  const long long r = t * WT + lane;
  for (int i = 0; i < dim; ++i) {
    double lo = r < n ? x[r * dim + i] : INFINITY, hi = r < n ? x[r * dim + i] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
      aabb[t * 2 * dim + i] = lo;
      aabb[t * 2 * dim + dim + i] = hi;
    }
  }
}

__global__ void aabb_super_kernel(double* aabb, int dim, long long tiles, long long supers) {
  const long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (s >= supers) return;
  const long long t = s * WS + lane;
  double* out = aabb + tiles * 2 * dim + s * 2 * dim;
  for (int i = 0; i < dim; ++i) {
    double lo = t < tiles ? aabb[t * 2 * dim + i] : INFINITY, hi = t < tiles ? aabb[t * 2 * dim + dim + i] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) out[i] = lo, out[dim + i] = hi;
  }
}

// Conservative decision margins.  The culls and the pair test evaluate an APPROXIMATE s (reciprocal
// multiply + FMA; relative error of a dozen ulp, << 1e-12, every term non-negative so nothing cancels):
//   approx s >= 1 + 1e-12   =>  the exactly-rounded reference s is > 1   (never stored)
//   approx s <  1 - 1e-12   =>  the exactly-rounded reference s is < 1   (always stored)
// and only the sliver in between is decided by the reference's own operation sequence (IEEE division,
// separately rounded multiply / add).  The pattern is therefore bit-exact while the geometry passes run
// three FP64 instructions per axis and pair.  VALUES are always computed from the exact sequence, by a
// separate lane-dense kernel over the stored entries (wendland_values_kernel).
#define FVGP_CULL_LIMIT (1.0 + 1e-12)
#define FVGP_SURE_LIMIT (1.0 - 1e-12)

// approximate s of the per-axis gaps between two boxes
template <int DIM>
__device__ __forceinline__ double box_gap_s(const double* lo1, const double* hi1, const double* lo2, const double* hi2,
                                            const double* rinv) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const double g = fmax(0.0, fmax(lo2[i] - hi1[i], lo1[i] - hi2[i]));
    const double t = g * rinv[i];
    s = fma(t, t, s);
  }
  return s;
}

// The reference's own sequence (kernels.py:520-523): subtract, TRUE divide, square, add -- each rounded separately.
template <int DIM>
__device__ __forceinline__ double exact_s(const double* xr, const double* xc, const double* theta) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const double t = __ddiv_rn(__dsub_rn(xr[i], xc[i]), theta[i]);
    s = __dadd_rn(s, __dmul_rn(t, t));
  }
  return s;
}

// kernels.py:524-528 on an exactly reproduced s < 1.
__device__ __forceinline__ double wendland_value(double s, double amp) {
  const double d = __dsqrt_rn(s);
  const double u = 1.0 - d;
  const double u2 = u * u, u4 = u2 * u2;
  const double d2 = d * d;
  const double poly = ((32.0 * (d2 * d) + 25.0 * d2) + 8.0 * d) + 1.0;
  return (amp * (u4 * u4)) * poly;
}

// One warp per UNIT = (32-row tile, chunk of the column super tiles).  Splitting the column range into up to 32
// chunks bounds the work of one warp: a row tile whose 32 consecutive points straddle a jump of the
// space-filling order has a huge bounding box and would otherwise test a large part of the matrix alone (ncu:
// SMs busy 1/3 of the kernel's duration, the rest was the tail of a few such warps).  Units without a surviving
// super tile cost one box test.  Lanes hold the 32 columns of the current column tile; lane r also keeps the
// running entry count (count pass) / write cursor (index pass) of row r in a register, so a hit costs one
// ballot, one shuffle and one predicated add -- no shared-memory traffic, no barriers.  Per-(chunk, row)
// counts go to p.chunk; wendland_chunk_prefix_kernel turns them into per-chunk offsets and row totals.
// STRICT (amp so small / large / non-finite that amp * (1-d)^8 * poly may round to 0 or inf): the hit
// decision additionally evaluates the value and applies the reference's `value != 0` test (np.nonzero,
// gp2Scale_covariance.py:147) for every candidate.
template <int DIM, bool FILL, bool STRICT>
__global__ void __launch_bounds__(W_WARPS * 32) wendland_csr_kernel(const WendlandParams p) {
  __shared__ double xs[W_WARPS][WT][DIM];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long unit = blockIdx.x * (long long)W_WARPS + warp;
  const long long tile = unit / p.n_chunks;
  const int chunk = (int)(unit - tile * p.n_chunks);
  if (tile >= p.tiles1) return;  // whole warp exits together; no CTA-wide barriers below
  const unsigned lt_mask = (1u << lane) - 1u;

  double theta[DIM], rinv[DIM], lo1[DIM], hi1[DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    theta[i] = p.theta[i];
    rinv[i] = 1.0 / fabs(p.theta[i]);
    lo1[i] = p.aabb1[tile * 2 * DIM + i];
    hi1[i] = p.aabb1[tile * 2 * DIM + DIM + i];
  }
  const long long r_mine = tile * WT + lane;
  const int rows_here = (int)min((long long)WT, p.n1 - tile * WT);
  long long cur = 0;  // lane r: entries of row r found by this unit (count) / write cursor of row r (fill)
  long long pairs = 0, row_tests = 0;
  bool staged = false;
  double xrow[DIM];   // this lane's row point (valid once staged)
#pragma unroll
  for (int i = 0; i < DIM; ++i) xrow[i] = 0.0;

  const double* tile_boxes = p.aabb2;
  const double* super_boxes = p.aabb2 + p.tiles2 * 2 * DIM;
  const long long s_begin = chunk * p.supers_per_chunk;
  const long long s_end = min(p.super2, s_begin + p.supers_per_chunk);

  for (long long s0 = s_begin; s0 < s_end; s0 += 32) {
    const long long sidx = s0 + lane;
    bool keep = false;
    if (sidx < s_end) {
      const double* bx = super_boxes + sidx * 2 * DIM;
      double lo2[DIM], hi2[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) lo2[i] = bx[i], hi2[i] = bx[DIM + i];
      keep = box_gap_s<DIM>(lo1, hi1, lo2, hi2, rinv) < FVGP_CULL_LIMIT;
    }
    unsigned smask = __ballot_sync(0xffffffffu, keep);
    if (smask != 0u && !staged) {  // first survivor: stage the row points, fetch the cursors
#pragma unroll
      for (int i = 0; i < DIM; ++i) {
        xrow[i] = r_mine < p.n1 ? p.x1[r_mine * DIM + i] : 0.0;
        xs[warp][lane][i] = xrow[i];
      }
      if (FILL && r_mine < p.n1) cur = p.indptr[r_mine] + p.chunk[(long long)chunk * p.n1 + r_mine];
      staged = true;
      __syncwarp();
    }
    while (smask) {
      const int sb = __ffs(smask) - 1;
      smask &= smask - 1;
      const long long t_idx = (s0 + sb) * WS + lane;
      bool keep_t = false;
      if (t_idx < p.tiles2) {
        const double* bx = tile_boxes + t_idx * 2 * DIM;
        double lo2[DIM], hi2[DIM];
#pragma unroll
        for (int i = 0; i < DIM; ++i) lo2[i] = bx[i], hi2[i] = bx[DIM + i];
        keep_t = box_gap_s<DIM>(lo1, hi1, lo2, hi2, rinv) < FVGP_CULL_LIMIT;
      }
      unsigned tmask = __ballot_sync(0xffffffffu, keep_t);
      pairs += __popc(tmask);
      while (tmask) {
        const int tb = __ffs(tmask) - 1;
        tmask &= tmask - 1;
        const long long ct = (s0 + sb) * WS + tb;
        // Per-ROW cull: lane r tests ITS row point against the column tile's box (broadcast loads of the box).  A
        // row tile's box inflated by the support radius is ~3x the radius wide at C4, so roughly half of its rows
        // are out of reach of any given surviving column tile; only the rows that pass enter the pair loop.
        unsigned rmask;
        {
          const double* bx = tile_boxes + ct * 2 * DIM;
          double gs = 0.0;
#pragma unroll
          for (int i = 0; i < DIM; ++i) {
            const double g = fmax(0.0, fmax(bx[i] - xrow[i], xrow[i] - bx[DIM + i]));
            const double t = g * rinv[i];
            gs = fma(t, t, gs);
          }
          rmask = __ballot_sync(0xffffffffu, lane < rows_here && gs < FVGP_CULL_LIMIT);
        }
        if (rmask == 0u) continue;
        row_tests += __popc(rmask);
        const long long c = ct * WT + lane;
        const bool c_ok = c < p.n2;
        double xc[DIM];
#pragma unroll
        for (int i = 0; i < DIM; ++i) xc[i] = c_ok ? p.x2[c * DIM + i] : 0.0;
        while (rmask) {
          // two rows per trip: two independent distance chains in flight
          const int ra = __ffs(rmask) - 1;
          rmask &= rmask - 1;
          const int rb = rmask ? __ffs(rmask) - 1 : -1;
          if (rb >= 0) rmask &= rmask - 1;
          double xa[DIM], xb[DIM];
#pragma unroll
          for (int i = 0; i < DIM; ++i) xa[i] = xs[warp][ra][i], xb[i] = xs[warp][rb >= 0 ? rb : ra][i];
          double sa = 0.0, sb2 = 0.0;
#pragma unroll
          for (int i = 0; i < DIM; ++i) {
            const double ta = (xa[i] - xc[i]) * rinv[i], tb2 = (xb[i] - xc[i]) * rinv[i];
            sa = fma(ta, ta, sa), sb2 = fma(tb2, tb2, sb2);
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int rr = half == 0 ? ra : rb;
            if (rr < 0) break;  // warp-uniform
            const double sv = half == 0 ? sa : sb2;
            const double* xr = half == 0 ? xa : xb;
            bool hit = c_ok && sv < FVGP_SURE_LIMIT;
            if (STRICT) {
              hit = false;
              if (c_ok && sv < FVGP_CULL_LIMIT) {
                const double s = exact_s<DIM>(xr, xc, theta);
                hit = s < 1.0 && wendland_value(s, p.amp) != 0.0;
              }
            } else if (c_ok && !hit && sv < FVGP_CULL_LIMIT) {
              hit = exact_s<DIM>(xr, xc, theta) < 1.0;  // the sliver: decided by the reference's sequence
            }
            const unsigned hmask = __ballot_sync(0xffffffffu, hit);
            if (hmask == 0u) continue;
            if (FILL) {
              const long long base = __shfl_sync(0xffffffffu, cur, rr);
              if (hit) p.indices[base + __popc(hmask & lt_mask)] = (int)c;
            }
            if (lane == rr) cur += __popc(hmask);
          }
        }
      }
    }
  }
  if (!FILL) {
    if (r_mine < p.n1) p.chunk[(long long)chunk * p.n1 + r_mine] = (int)cur;
    if (p.stats != nullptr && lane == 0 && pairs != 0) {   // stats[0]: tile pairs that reached the pair loop's door,
      atomicAdd((unsigned long long*)p.stats, (unsigned long long)pairs);            // stats[1]: candidate pairs tested
      atomicAdd((unsigned long long*)p.stats + 1, (unsigned long long)row_tests * 32ull);
    }
  }
}

// Per row: exclusive prefix of its per-chunk counts (in place) and the row total.
__global__ void wendland_chunk_prefix_kernel(int* chunk, long long n1, int n_chunks, long long* rowcount) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= n1) return;
  long long run = 0;
  for (int c = 0; c < n_chunks; ++c) {
    const int v = chunk[(long long)c * n1 + r];
    chunk[(long long)c * n1 + r] = (int)run;
    run += v;
  }
  rowcount[r] = run;
}

// Values of the stored entries, lane-dense: one warp per row walks the row's column indices.  The exact
// operation sequence of the reference gives s (bit-identical to numpy), the value follows kernels.py:524-528;
// the noise diagonal of K + diag(V) (gp_kv.py:655-661) is fused.
template <int DIM>
__global__ void __launch_bounds__(256) wendland_values_kernel(const WendlandParams p) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  double theta[DIM];
#pragma unroll
  for (int i = 0; i < DIM; ++i) theta[i] = p.theta[i];
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.n1; row += warps) {
    const long long b = p.indptr[row], e = p.indptr[row + 1];
    double xr[DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i) xr[i] = p.x1[row * DIM + i];
    for (long long k = b + lane; k < e; k += 32) {
      const long long c = p.indices[k];
      double xc[DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i) xc[i] = p.x2[c * DIM + i];
      double v = wendland_value(exact_s<DIM>(xr, xc, theta), p.amp);
      if (p.noise != nullptr && c == row + p.row0) v += p.noise[row];
      p.data[k] = v;
    }
  }
}

// ---------------------------------------------------------------------------- exclusive scan (int64)
constexpr int SCAN_CHUNK = 2048;

__global__ void __launch_bounds__(256) scan_block_sums_kernel(const long long* __restrict__ in, long long n,
                                                              long long* sums) {
  __shared__ long long red[256];
  const long long base = blockIdx.x * (long long)SCAN_CHUNK;
  long long s = 0;
  for (int k = threadIdx.x; k < SCAN_CHUNK; k += 256)
    if (base + k < n) s += in[base + k];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[blockIdx.x] = red[0];
}

__global__ void scan_sums_kernel(long long* sums, long long nblk) {  // single thread: nblk is small
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long run = 0;
    for (long long i = 0; i < nblk; ++i) {
      const long long v = sums[i];
      sums[i] = run;
      run += v;
    }
    sums[nblk] = run;
  }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(const long long* __restrict__ in, long long n,
                                                         const long long* __restrict__ sums, long long* out,
                                                         long long nblk) {
  __shared__ long long part[256];
  const long long base = blockIdx.x * (long long)SCAN_CHUNK + threadIdx.x * 8;
  long long v[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  part[threadIdx.x] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over 256 thread sums
  for (int o = 1; o < 256; o <<= 1) {
    long long add = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += add;
    __syncthreads();
  }
  long long run = sums[blockIdx.x] + part[threadIdx.x] - s;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (blockIdx.x == nblk - 1 && threadIdx.x == 0) out[n] = sums[nblk];
}

// ---------------------------------------------------------------------------- SpMV and Krylov kernels
constexpr int KR_THREADS = 256;

struct KrylovScalars {
  double rho, rho_old, pq, rr, atol, alpha, beta, bnorm2;
  int done, iters, calls, pad;
  unsigned counter[4];
};

// LPR lanes per row (32 / LPR rows per warp), U strips of LPR entries of every row in flight at once: the
// (val, idx) loads of all strips are issued before the first dependent gather of x.  SpMV on the gp2Scale
// matrix (~100 entries per row) is latency bound, not HBM bound, until enough bytes are in flight (ncu: 72 % of
// the warp cycles are long-scoreboard stalls): what counts is rows-in-flight x occupancy, so the variants are
// register-capped.  Returns the row sum in all LPR lanes of the group; groups with row >= n return 0.
template <int LPR, int U>
__device__ __forceinline__ double csr_group_dot(const long long* __restrict__ indptr, const int* __restrict__ idx,
                                                const double* __restrict__ val, const double* __restrict__ x,
                                                long long row, long long n, int sub) {
  const bool valid = row < n;
  const long long b = valid ? indptr[row] : 0, e = valid ? indptr[row + 1] : 0;
  double acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = 0.0;
  for (long long k = b + sub; k < e; k += LPR * U) {
    double v[U];
    int j[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool on = k + u * LPR < e;
      v[u] = on ? val[k + u * LPR] : 0.0;
      j[u] = on ? idx[k + u * LPR] : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double xv = j[u] >= 0 ? x[j[u]] : 0.0;
      acc[u] = fma(v[u], xv, acc[u]);
    }
  }
  double s = acc[0];
#pragma unroll
  for (int u = 1; u < U; ++u) s += acc[u];
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

// Row loop shared by the three kernels that multiply by the matrix: f(row, sum) runs in the group's lane 0.
template <int LPR, int U, typename F>
__device__ __forceinline__ void spmv_rows(long long n, const long long* __restrict__ indptr,
                                          const int* __restrict__ idx, const double* __restrict__ val,
                                          const double* __restrict__ x, F f) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, grp = lane / LPR, sub = lane % LPR;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; base < n;
       base += warps * RPW) {
    const long long row = base + grp;
    const double s = csr_group_dot<LPR, U>(indptr, idx, val, x, row, n, sub);
    if (sub == 0 && row < n) f(row, s);
  }
}

template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(KR_THREADS, MINB) spmv_kernel(long long n, const long long* __restrict__ indptr,
                                                                const int* __restrict__ idx,
                                                                const double* __restrict__ val,
                                                                const double* __restrict__ x, double* __restrict__ y) {
  spmv_rows<LPR, U>(n, indptr, idx, val, x, [&](long long row, double s) { y[row] = s; });
}

// Deterministic grid-wide sums: every CTA publishes its partial, the last CTA to arrive adds
// them in index order.  Returns true in the finishing CTA (all its threads), totals in out[].
template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double* partials, unsigned* counter, double (&out)[NV]) {
  __shared__ double red[32];
  __shared__ bool last;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    const double s = block_sum(v[q], red);
    if (threadIdx.x == 0) partials[(size_t)q * gridDim.x + blockIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!last) return false;
  __threadfence();
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += partials[(size_t)q * gridDim.x + i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) red[0] = s;
    __syncthreads();
    out[q] = red[0];
    __syncthreads();
  }
  return true;
}

// r = b - A x ; bnorm2 = b.b ; sets the absolute tolerance.
template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(KR_THREADS, MINB) pcg_init_kernel(long long n, const long long* __restrict__ indptr,
                                                                    const int* __restrict__ idx,
                                                                    const double* __restrict__ val,
                                                                    const double* __restrict__ b,
                                                                    const double* __restrict__ x, double* __restrict__ r,
                                                                    double rtol, KrylovScalars* sc, double* partials) {
  double v[2] = {0.0, 0.0};
  spmv_rows<LPR, U>(n, indptr, idx, val, x, [&](long long row, double ax) {
    const double bi = b[row], ri = bi - ax;
    r[row] = ri;
    v[0] += bi * bi;
    v[1] += ri * ri;
  });
  double tot[2];
  if (grid_sum<2>(v, partials, &sc->counter[0], tot) && threadIdx.x == 0) {
    sc->bnorm2 = tot[0];
    sc->rr = tot[1];
    sc->atol = rtol * sqrt(tot[0]);
    sc->rho = 0.0, sc->rho_old = 0.0, sc->iters = 0, sc->calls = 0, sc->alpha = 0.0;
    sc->done = (tot[0] == 0.0) ? 1 : 0;  // scipy returns at once for b = 0
  }
}

// Head of iteration k, fused with the tail of iteration k-1 (one pass over the vectors):
//   x += alpha p ; r -= alpha q              (alpha of iteration k-1; 0 on the first call)
//   ||r|| < atol ?  -> done                  (scipy checks at the TOP of every iteration)
//   z = M r  (block-Jacobi: 32x32 symmetric inverse blocks; M = I when blocks == NULL)
//   rho = r.z ; beta = rho / rho_old
__global__ void __launch_bounds__(KR_THREADS) pcg_head_kernel(long long n, const double* __restrict__ blocks,
                                                              const double* __restrict__ p, const double* __restrict__ q,
                                                              double* __restrict__ x, double* __restrict__ r,
                                                              double* __restrict__ z, KrylovScalars* sc,
                                                              double* partials, int maxiter) {
  if (sc->done) return;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nblk = (n + 31) / 32;
  const double alpha = sc->alpha;
  const bool update = sc->calls > 0;
  double v[2] = {0.0, 0.0};
  for (long long blk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < nblk; blk += warps) {
    const long long i = blk * 32 + lane;
    double ri = 0.0;
    if (i < n) {
      ri = r[i];
      if (update) {
        x[i] = fma(alpha, p[i], x[i]);
        ri = fma(-alpha, q[i], ri);
        r[i] = ri;
      }
    }
    double zi = ri;
    if (blocks != nullptr) {
      const double* B = blocks + blk * 1024;
      zi = 0.0;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) zi = fma(B[j * 32 + lane], __shfl_sync(0xffffffffu, ri, j), zi);
    }
    if (i < n) {
      z[i] = zi;
      v[0] += ri * ri;
      v[1] += ri * zi;
    }
  }
  double tot[2];
  if (grid_sum<2>(v, partials, &sc->counter[1], tot) && threadIdx.x == 0) {
    sc->rr = tot[0];
    sc->iters = sc->calls;  // completed x / r updates
    sc->calls += 1;
    if (sqrt(tot[0]) < sc->atol) sc->done = 1;
    else if (sc->iters >= maxiter) sc->done = 2;
    sc->rho_old = sc->rho;
    sc->rho = tot[1];
    sc->beta = sc->iters > 0 ? tot[1] / sc->rho_old : 0.0;
  }
}

__global__ void __launch_bounds__(KR_THREADS) pcg_update_p_kernel(long long n, const double* __restrict__ z,
                                                                  double* __restrict__ p, const KrylovScalars* sc) {
  if (sc->done) return;
  const double beta = sc->beta;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = z[i] + beta * p[i];
}

// q = A p ; pq = p.q ; alpha = rho / pq
template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(KR_THREADS, MINB) pcg_spmv_kernel(long long n, const long long* __restrict__ indptr,
                                                                    const int* __restrict__ idx,
                                                                    const double* __restrict__ val,
                                                                    const double* __restrict__ p, double* __restrict__ q,
                                                                    KrylovScalars* sc, double* partials) {
  if (sc->done) return;
  double v[1] = {0.0};
  spmv_rows<LPR, U>(n, indptr, idx, val, p, [&](long long row, double s) {
    q[row] = s;
    v[0] += s * p[row];
  });
  double tot[1];
  if (grid_sum<1>(v, partials, &sc->counter[2], tot) && threadIdx.x == 0) {
    sc->pq = tot[0];
    sc->alpha = sc->rho / tot[0];
  }
}

// ---------------------------------------------------------------------------- row-sharded PCG (multi-GPU)
// Rank r owns rows [row0, row0 + nrows) of the matrix and of x, r, z, q; the search direction p is kept in full on
// every rank (the slab SpMV gathers from all of it).  One iteration = the single-GPU iteration with three
// collectives spliced in (fvgp_pcg_sharded): the local partial sums of (r.r, r.z) and of p.q are all-reduced in
// place in sc->loc, p's slabs are all-gathered after the update.  The derived scalars (rho, alpha, beta, the
// convergence flag) are computed redundantly on every rank from the same reduced values by a one-thread kernel, so
// every rank takes the same branch and issues the same sequence of collectives.
struct ShardScalars {
  double loc[4];  // [0] r.r (b.b at init)  [1] r.z (r.r at init)  [2] p.q      -- all-reduced in place
  double rho, rho_old, atol, alpha, beta, bnorm2, rr;
  int done, iters, calls, pad;
  unsigned counter[4];
};

template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(KR_THREADS, MINB) pcgs_init_kernel(long long nrows, const long long* __restrict__ indptr,
                                                                     const int* __restrict__ idx,
                                                                     const double* __restrict__ val,
                                                                     const double* __restrict__ b_slab,
                                                                     const double* __restrict__ x_full,
                                                                     double* __restrict__ r, ShardScalars* sc,
                                                                     double* partials) {
  double v[2] = {0.0, 0.0};
  spmv_rows<LPR, U>(nrows, indptr, idx, val, x_full, [&](long long row, double ax) {
    const double bi = b_slab[row], ri = bi - ax;
    r[row] = ri;
    v[0] += bi * bi;
    v[1] += ri * ri;
  });
  double tot[2];
  if (grid_sum<2>(v, partials, &sc->counter[0], tot) && threadIdx.x == 0) sc->loc[0] = tot[0], sc->loc[1] = tot[1];
}

// which = 0 after the init reduction, 1 after the head reduction, 2 after the p.q reduction
__global__ void pcgs_scalar_kernel(ShardScalars* sc, int which, double rtol, int maxiter) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (which == 0) {
    sc->bnorm2 = sc->loc[0];
    sc->rr = sc->loc[1];
    sc->atol = rtol * sqrt(sc->loc[0]);
    sc->rho = 0.0, sc->rho_old = 0.0, sc->iters = 0, sc->calls = 0, sc->alpha = 0.0, sc->beta = 0.0;
    sc->done = (sc->loc[0] == 0.0) ? 1 : 0;  // scipy returns at once for b = 0
    return;
  }
  if (sc->done) return;
  if (which == 1) {
    sc->rr = sc->loc[0];
    sc->iters = sc->calls;  // completed x / r updates
    sc->calls += 1;
    if (sqrt(sc->loc[0]) < sc->atol) sc->done = 1;
    else if (sc->iters >= maxiter) sc->done = 2;
    sc->rho_old = sc->rho;
    sc->rho = sc->loc[1];
    sc->beta = sc->iters > 0 ? sc->loc[1] / sc->rho_old : 0.0;
  } else {
    sc->alpha = sc->rho / sc->loc[2];
  }
}

// x += alpha p ; r -= alpha q ; z = M r ; local sums of r.r and r.z   (all pointers are slab-local)
__global__ void __launch_bounds__(KR_THREADS) pcgs_head_kernel(long long nrows, const double* __restrict__ blocks,
                                                               const double* __restrict__ p, const double* __restrict__ q,
                                                               double* __restrict__ x, double* __restrict__ r,
                                                               double* __restrict__ z, ShardScalars* sc,
                                                               double* partials) {
  if (sc->done) return;
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long nblk = (nrows + 31) / 32;
  const double alpha = sc->alpha;
  const bool update = sc->calls > 0;
  double v[2] = {0.0, 0.0};
  for (long long blk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < nblk; blk += warps) {
    const long long i = blk * 32 + lane;
    double ri = 0.0;
    if (i < nrows) {
      ri = r[i];
      if (update) {
        x[i] = fma(alpha, p[i], x[i]);
        ri = fma(-alpha, q[i], ri);
        r[i] = ri;
      }
    }
    double zi = ri;
    if (blocks != nullptr) {
      const double* B = blocks + blk * 1024;
      zi = 0.0;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) zi = fma(B[j * 32 + lane], __shfl_sync(0xffffffffu, ri, j), zi);
    }
    if (i < nrows) {
      z[i] = zi;
      v[0] += ri * ri;
      v[1] += ri * zi;
    }
  }
  double tot[2];
  if (grid_sum<2>(v, partials, &sc->counter[1], tot) && threadIdx.x == 0) sc->loc[0] = tot[0], sc->loc[1] = tot[1];
}

__global__ void __launch_bounds__(KR_THREADS) pcgs_update_p_kernel(long long nrows, const double* __restrict__ z,
                                                                   double* __restrict__ p_slab, const ShardScalars* sc) {
  if (sc->done) return;
  const double beta = sc->beta;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += (long long)gridDim.x * blockDim.x)
    p_slab[i] = z[i] + beta * p_slab[i];
}

// q = A_slab p ; local p.q
template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(KR_THREADS, MINB) pcgs_spmv_kernel(long long nrows, const long long* __restrict__ indptr,
                                                                     const int* __restrict__ idx,
                                                                     const double* __restrict__ val,
                                                                     const double* __restrict__ p_full,
                                                                     const double* __restrict__ p_slab,
                                                                     double* __restrict__ q, ShardScalars* sc,
                                                                     double* partials) {
  if (sc->done) return;
  double v[1] = {0.0};
  spmv_rows<LPR, U>(nrows, indptr, idx, val, p_full, [&](long long row, double s) {
    q[row] = s;
    v[0] += s * p_slab[row];
  });
  double tot[1];
  if (grid_sum<1>(v, partials, &sc->counter[2], tot) && threadIdx.x == 0) sc->loc[2] = tot[0];
}

// Dense 32x32 diagonal blocks of the CSR matrix, inverted in shared memory (Gauss-Jordan, SPD).
__global__ void __launch_bounds__(128) bjacobi_build_kernel(long long n, const long long* __restrict__ indptr,
                                                            const int* __restrict__ idx,
                                                            const double* __restrict__ val, double* blocks) {
  __shared__ double Bs[4][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long blk = blockIdx.x * 4ll + warp;
  const long long nblk = (n + 31) / 32;
  if (blk >= nblk) return;
  double(*B)[33] = Bs[warp];
  for (int j = 0; j < 32; ++j) B[lane][j] = (j == lane) ? 1.0 : 0.0;  // identity padding for the ragged tail
  __syncwarp();
  const long long row = blk * 32 + lane, c0 = blk * 32;
  if (row < n) {
    bool any = false;
    for (long long k = indptr[row]; k < indptr[row + 1]; ++k) {
      const long long c = idx[k];
      if (c >= c0 && c < c0 + 32) {
        if (!any) B[lane][lane] = 0.0, any = true;
        B[lane][c - c0] += val[k];
      }
    }
  }
  __syncwarp();
  for (int k = 0; k < 32; ++k) {
    const double piv = 1.0 / B[k][k];
    __syncwarp();
    const double f = B[lane][k];
    if (lane == k) {
      for (int j = 0; j < 32; ++j) B[k][j] = (j == k) ? piv : B[k][j] * piv;
    }
    __syncwarp();
    if (lane != k) {
      for (int j = 0; j < 32; ++j) B[lane][j] = (j == k) ? -f * piv : B[lane][j] - f * B[k][j];
    }
    __syncwarp();
  }
  double* out = blocks + blk * 1024;
  for (int j = 0; j < 32; ++j) out[j * 32 + lane] = 0.5 * (B[j][lane] + B[lane][j]);  // symmetrised
}

// ---------------------------------------------------------------------------- Lanczos (SLQ), NB probes at once
// The probes of one batch advance in lock step as the columns of n x NB row-major blocks, so the matrix is
// streamed once per Lanczos step for NB probes (SpMM) and each gather of a neighbour touches NB contiguous
// doubles.  Vectors are kept UNNORMALISED (u_j, with v_j = s_j u_j and s_j = 1 / beta_j held on the
// device), which removes the normalise-and-shift pass: three buffers rotate on the host side.
//   step j:  w = s_j A u_j - beta_j s_{j-1} u_{j-1} ;  alpha_j = w . v_j          (lanczos_spmm_kernel)
//            w -= alpha_j v_j ;  beta_{j+1} = ||w|| ;  u_{j+1} = w                 (lanczos_axpy_kernel)
constexpr int LZ_MAXB = 16;

struct LanczosScalars {
  double s_cur[LZ_MAXB], s_prev[LZ_MAXB], beta[LZ_MAXB], alpha[LZ_MAXB];
  unsigned counter[4];
};

__device__ __forceinline__ double rademacher(unsigned long long seed, unsigned long long probe, unsigned long long i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (probe * 0x100000001B3ull + i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (z & 1ull) ? 1.0 : -1.0;
}

template <int NB>
__global__ void lanczos_start_kernel(long long n, unsigned long long seed, unsigned long long probe0, double* u,
                                     double* uprev, LanczosScalars* sc) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * NB;
       i += (long long)gridDim.x * blockDim.x) {
    u[i] = rademacher(seed, probe0 + (unsigned long long)(i % NB), (unsigned long long)(i / NB));
    uprev[i] = 0.0;
  }
  if (blockIdx.x == 0 && threadIdx.x < NB) {
    sc->s_cur[threadIdx.x] = rsqrt((double)n);
    sc->s_prev[threadIdx.x] = 0.0;
    sc->beta[threadIdx.x] = 0.0;
  }
}

// Sum NB per-lane accumulators over the warp, halving the number of live values per exchange; afterwards
// lane l holds the total of column l / (32 / NB).  NB + log2(32 / NB) - 1 shuffles instead of 5 NB.
template <int NB>
__device__ __forceinline__ double warp_sum_columns(double (&acc)[NB], int lane) {
  int off = 16;
#pragma unroll
  for (int w = NB / 2; w >= 1; w >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const double send = upper ? acc[i] : acc[i + w];
      const double keep = upper ? acc[i + w] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  double v = acc[0];
  for (; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Deterministic grid-wide sum of NB column values held by the first NB threads' slots of every warp.
template <int NB>
__device__ __forceinline__ bool grid_sum_columns(double mine, bool owner, int col, double* partials, unsigned* counter,
                                                 double* out /* shared, NB */) {
  __shared__ double red[KR_THREADS / 32][LZ_MAXB];
  __shared__ bool last;
  const int wid = threadIdx.x >> 5;
  if (owner) red[wid][col] = mine;
  __syncthreads();
  if (threadIdx.x < NB) {
    double s = 0.0;
    for (int w = 0; w < KR_THREADS / 32; ++w) s += red[w][threadIdx.x];
    partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (!last) return false;
  __threadfence();
  // finishing CTA: warp c sums column c in index order (KR_THREADS / 32 = 8 warps; NB <= 16 -> two rounds)
  const int lane = threadIdx.x & 31;
  for (int c = wid; c < NB; c += KR_THREADS / 32) {
    double s = 0.0;
    for (unsigned i = lane; i < gridDim.x; i += 32) s += partials[(size_t)c * gridDim.x + i];
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
  }
  __syncthreads();
  return true;
}

template <int NB>
__global__ void __launch_bounds__(KR_THREADS) lanczos_spmm_kernel(long long n, const long long* __restrict__ indptr,
                                                                  const int* __restrict__ idx,
                                                                  const double* __restrict__ val,
                                                                  const double* __restrict__ u,
                                                                  const double* __restrict__ uprev,
                                                                  double* __restrict__ w, LanczosScalars* sc,
                                                                  double* partials) {
  __shared__ double tot[LZ_MAXB];
  const int lane = threadIdx.x & 31;
  constexpr int LPC = 32 / NB;  // lanes per column after the reduction
  const int col = lane / LPC;
  const bool owner = (lane % LPC) == 0;
  const double s_cur = sc->s_cur[col], bs_prev = sc->beta[col] * sc->s_prev[col];
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  double asum = 0.0;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const long long b = indptr[row], e = indptr[row + 1];
    double acc[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) acc[c] = 0.0;
    for (long long k = b + lane; k < e; k += 64) {
      const bool p1 = k + 32 < e;
      const double v0 = val[k];
      const long long j0 = idx[k];
      const double v1 = p1 ? val[k + 32] : 0.0;
      const long long j1 = p1 ? idx[k + 32] : j0;
      if constexpr (NB == 1) {
        acc[0] = fma(v0, u[j0], acc[0]);
        acc[0] = fma(v1, u[j1], acc[0]);
      } else {
        const double2* g0 = reinterpret_cast<const double2*>(u + j0 * NB);
        const double2* g1 = reinterpret_cast<const double2*>(u + j1 * NB);
        double2 a[NB / 2], bb[NB / 2];
#pragma unroll
        for (int c = 0; c < NB / 2; ++c) a[c] = g0[c];
#pragma unroll
        for (int c = 0; c < NB / 2; ++c) bb[c] = g1[c];
#pragma unroll
        for (int c = 0; c < NB / 2; ++c) {
          acc[2 * c] = fma(v0, a[c].x, acc[2 * c]);
          acc[2 * c + 1] = fma(v0, a[c].y, acc[2 * c + 1]);
          acc[2 * c] = fma(v1, bb[c].x, acc[2 * c]);
          acc[2 * c + 1] = fma(v1, bb[c].y, acc[2 * c + 1]);
        }
      }
    }
    const double au = warp_sum_columns<NB>(acc, lane);
    if (owner) {
      const double wi = s_cur * au - bs_prev * uprev[row * NB + col];
      w[row * NB + col] = wi;
      asum = fma(wi, s_cur * u[row * NB + col], asum);
    }
  }
  if (grid_sum_columns<NB>(asum, owner, col, partials, &sc->counter[0], tot) && threadIdx.x < NB)
    sc->alpha[threadIdx.x] = tot[threadIdx.x];
}

// Same step with LANES = COLUMNS (NB >= 4): a group of NB lanes owns one row (32 / NB rows per warp), lane c
// accumulates probe c.  The group streams the row NB entries at a time (lane u loads entry k + u, coalesced),
// broadcasts (value, column) with shuffles and gathers u[column * NB + c] -- NB contiguous doubles per entry
// across the group.  No cross-lane reduction, ~36 registers (full occupancy) and 32 / NB rows x NB gathers in
// flight per warp; the per-lane-entry variant above needs NB accumulators + NB gathered doubles per lane
// (64 registers at NB = 8, half occupancy) and was latency bound (ncu: 17 % issue-active, 13 % DRAM).
template <int NB>
__global__ void __launch_bounds__(KR_THREADS, (NB <= 4 ? 6 : NB == 8 ? 4 : 3)) lanczos_spmm_cols_kernel(
    long long n, const long long* __restrict__ indptr, const int* __restrict__ idx, const double* __restrict__ val,
    const double* __restrict__ u, const double* __restrict__ uprev, double* __restrict__ w, LanczosScalars* sc,
    double* partials) {
  __shared__ double tot[LZ_MAXB];
  constexpr int RPW = 32 / NB;
  const int lane = threadIdx.x & 31, grp = lane / NB, c = lane % NB;
  const double s_cur = sc->s_cur[c], bs_prev = sc->beta[c] * sc->s_prev[c];
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  double asum = 0.0;
  for (long long base = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * RPW; base < n;
       base += warps * RPW) {
    const long long row = base + grp;
    const bool valid = row < n;
    const long long b = valid ? indptr[row] : 0, e = valid ? indptr[row + 1] : 0;
    // every group of the warp runs the same number of rounds (shuffles need all lanes)
    long long len = e - b;
#pragma unroll
    for (int o = NB; o < 32; o <<= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    double acc0 = 0.0, acc1 = 0.0;
    for (long long k0 = 0; k0 < len; k0 += NB) {
      const long long ka = b + k0 + c;
      const double va = ka < e ? val[ka] : 0.0;
      const int ja = ka < e ? idx[ka] : -1;
      double xa[NB];
#pragma unroll
      for (int t = 0; t < NB; ++t) {  // all NB gathers of the round are issued before the first use
        const int j1 = __shfl_sync(0xffffffffu, ja, grp * NB + t);
        xa[t] = j1 >= 0 ? u[(long long)j1 * NB + c] : 0.0;
      }
#pragma unroll
      for (int t = 0; t < NB; t += 2) {
        acc0 = fma(__shfl_sync(0xffffffffu, va, grp * NB + t), xa[t], acc0);
        acc1 = fma(__shfl_sync(0xffffffffu, va, grp * NB + t + 1), xa[t + 1], acc1);
      }
    }
    if (valid) {
      const double wi = s_cur * (acc0 + acc1) - bs_prev * uprev[row * NB + c];
      w[row * NB + c] = wi;
      asum = fma(wi, s_cur * u[row * NB + c], asum);
    }
  }
#pragma unroll
  for (int o = 16; o >= NB; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
  if (grid_sum_columns<NB>(asum, lane < NB, lane % NB, partials, &sc->counter[0], tot) && threadIdx.x < NB)
    sc->alpha[threadIdx.x] = tot[threadIdx.x];
}

// w -= alpha v ; beta_next = ||w|| per column; rotates the scale factors and records (alpha_j, beta_{j+1}).
template <int NB>
__global__ void __launch_bounds__(KR_THREADS) lanczos_axpy_kernel(long long n, const double* __restrict__ u,
                                                                  double* __restrict__ w, LanczosScalars* sc,
                                                                  double* partials, double* out_alpha,
                                                                  double* out_beta) {
  __shared__ double tot[LZ_MAXB];
  // thread t always works on column t % NB (the grid stride is a multiple of NB)
  const int col = threadIdx.x % NB;
  const double as = sc->alpha[col] * sc->s_cur[col];
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * NB;
       i += (long long)gridDim.x * blockDim.x) {
    const double wi = fma(-as, u[i], w[i]);
    w[i] = wi;
    acc = fma(wi, wi, acc);
  }
  // fold the 32 / NB lanes of a warp that share a column, then the deterministic grid sum
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= NB; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (grid_sum_columns<NB>(acc, lane < NB, lane % NB, partials, &sc->counter[1], tot) && threadIdx.x < NB) {
    const int c = threadIdx.x;
    const double beta = sqrt(tot[c]);
    out_alpha[c] = sc->alpha[c];
    out_beta[c] = beta;
    sc->s_prev[c] = sc->s_cur[c];
    sc->s_cur[c] = beta > 0.0 ? 1.0 / beta : 0.0;
    sc->beta[c] = beta;
  }
}

static inline unsigned krylov_grid() { return (unsigned)sm_count() * 8u; }  // 8 x 256 threads = full occupancy

// SpMV variant (lanes per row, strips in flight, minimum resident CTAs per SM = register cap):
//   0: 32 / 1 / 8   one row per warp, the classic CSR-vector loop
//   1:  8 / 4 / 8   four rows per warp, <= 32 registers
//   2:  8 / 8 / 5   four rows per warp, deeper prefetch, 48 registers
//   3: 16 / 4 / 8   two rows per warp
// FVGP_SPMV_VARIANT overrides the default (chosen from the B200 A/B in profiles/).
static int spmv_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FVGP_SPMV_VARIANT");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 3) v = 1;
  }
  return v;
}
#define FVGP_SPMV_DISPATCH(KERNEL, ...)                                         \
  switch (spmv_variant()) {                                                     \
    case 0: launch(KERNEL<32, 1, 8>, __VA_ARGS__); break;                       \
    case 2: launch(KERNEL<8, 8, 5>, __VA_ARGS__); break;                        \
    case 3: launch(KERNEL<16, 4, 8>, __VA_ARGS__); break;                       \
    default: launch(KERNEL<8, 4, 8>, __VA_ARGS__); break;                       \
  }

// Column super tiles are cut into at most 32 chunks of a multiple of 32 super tiles each.
static void wendland_chunking(int64_t n2, long long& supers_per_chunk, int& n_chunks) {
  const long long tiles2 = (n2 + WT - 1) / WT, super2 = (tiles2 + WS - 1) / WS;
  supers_per_chunk = 32 * std::max<long long>(1, (super2 + 1023) / 1024);
  n_chunks = (int)std::max<long long>(1, (super2 + supers_per_chunk - 1) / supers_per_chunk);
}

// amp * (1-d)^8 * poly with (1-d)^8 >= 2^-424 and 1 <= poly <= 66 is a normal non-zero double for every
// stored pair iff amp is moderately scaled; otherwise the reference's `value != 0` test must be replayed.
static inline bool wendland_needs_strict(double amp) {
  const double a = fabs(amp);
  return !(a >= 1e-150 && a <= 1e150);
}

template <bool FILL>
static int launch_wendland(const WendlandParams& p, cudaStream_t st) {
  const unsigned grid = (unsigned)((p.tiles1 * p.n_chunks + W_WARPS - 1) / W_WARPS);
  const bool strict = wendland_needs_strict(p.amp);
#define FVGP_WL(D)                                                                            \
  case D:                                                                                     \
    if (strict) launch(wendland_csr_kernel<D, FILL, true>, grid, W_WARPS * 32, 0, st, p);     \
    else launch(wendland_csr_kernel<D, FILL, false>, grid, W_WARPS * 32, 0, st, p);           \
    break;
  switch (p.dim) {
    FVGP_WL(1) FVGP_WL(2) FVGP_WL(3) FVGP_WL(4) FVGP_WL(5) FVGP_WL(6)
    default: FVGP_REQUIRE(!"gp2Scale Wendland supports 1..6 input dimensions");
  }
#undef FVGP_WL
  FVGP_LAUNCH_OK();
  return 0;
}

static int launch_wendland_values(const WendlandParams& p, cudaStream_t st) {
  const long long want = (p.n1 * 32 + 255) / 256;
  const unsigned grid = (unsigned)(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
  switch (p.dim) {
    case 1: launch(wendland_values_kernel<1>, grid, 256, 0, st, p); break;
    case 2: launch(wendland_values_kernel<2>, grid, 256, 0, st, p); break;
    case 3: launch(wendland_values_kernel<3>, grid, 256, 0, st, p); break;
    case 4: launch(wendland_values_kernel<4>, grid, 256, 0, st, p); break;
    case 5: launch(wendland_values_kernel<5>, grid, 256, 0, st, p); break;
    case 6: launch(wendland_values_kernel<6>, grid, 256, 0, st, p); break;
    default: FVGP_REQUIRE(!"gp2Scale Wendland supports 1..6 input dimensions");
  }
  FVGP_LAUNCH_OK();
  return 0;
}

static int fill_wendland_params(WendlandParams& p, const double* d_x1, int64_t n1, const double* d_aabb1,
                                const double* d_x2, int64_t n2, const double* d_aabb2, int dim,
                                const double* h_theta) {
  FVGP_REQUIRE(dim >= 1 && dim <= 6 && n1 >= 0 && n2 >= 0 && n2 < (1ll << 31));
  p.x1 = d_x1, p.x2 = d_x2, p.aabb1 = d_aabb1, p.aabb2 = d_aabb2;
  p.n1 = n1, p.n2 = n2, p.dim = dim, p.amp = h_theta[0];
  p.tiles1 = (n1 + WT - 1) / WT, p.tiles2 = (n2 + WT - 1) / WT, p.super2 = (p.tiles2 + WS - 1) / WS;
  wendland_chunking(n2, p.supers_per_chunk, p.n_chunks);
  p.chunk = nullptr;
  p.row0 = 0;
  for (int i = 0; i < kMaxDim; ++i) p.theta[i] = i < dim ? h_theta[1 + i] : 1.0;
  p.indptr = nullptr, p.noise = nullptr, p.rowcount = nullptr, p.stats = nullptr, p.indices = nullptr, p.data = nullptr;
  return 0;
}


}  // namespace fvgp

using namespace fvgp;

extern "C" {

int64_t fvgp_wendland_aabb_len(int64_t n, int dim) {
  const int64_t tiles = (n + WT - 1) / WT, supers = (tiles + WS - 1) / WS;
  return (tiles + supers) * 2 * dim;
}

int64_t fvgp_wendland_chunk_len(int64_t n1, int64_t n2) {
  long long spc;
  int nc;
  wendland_chunking(n2, spc, nc);
  return (int64_t)nc * std::max<int64_t>(n1, 1);
}

int fvgp_wendland_aabb(const double* d_x, int64_t n, int dim, double* d_aabb, void* stream) {
  FVGP_REQUIRE(dim >= 1 && dim <= kMaxDim);
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = (n + WT - 1) / WT, supers = (tiles + WS - 1) / WS;
  launch(aabb_tiles_kernel, (unsigned)((tiles * 32 + 255) / 256), 256, 0, st, d_x, n, dim, d_aabb, tiles);
  FVGP_LAUNCH_OK();
  launch(aabb_super_kernel, (unsigned)((supers * 32 + 255) / 256), 256, 0, st, d_aabb, dim, tiles, supers);
  FVGP_LAUNCH_OK();
  return 0;
}

int fvgp_wendland_csr_count(const double* d_x1, int64_t n1, const double* d_aabb1, const double* d_x2, int64_t n2,
                            const double* d_aabb2, int dim, const double* h_theta, int64_t* d_rowcount,
                            int32_t* d_chunk, int64_t* d_stats, void* stream) {
  WendlandParams p;
  int r = fill_wendland_params(p, d_x1, n1, d_aabb1, d_x2, n2, d_aabb2, dim, h_theta);
  if (r != 0) return r;
  if (n1 == 0) return 0;
  p.rowcount = (long long*)d_rowcount, p.stats = (long long*)d_stats, p.chunk = d_chunk;
  r = launch_wendland<false>(p, (cudaStream_t)stream);
  if (r != 0) return r;
  launch(wendland_chunk_prefix_kernel, (unsigned)((n1 + 255) / 256), 256, 0, (cudaStream_t)stream, d_chunk, (long long)n1,
         p.n_chunks, (long long*)d_rowcount);
  FVGP_LAUNCH_OK();
  return 0;
}

int fvgp_wendland_csr_fill(const double* d_x1, int64_t n1, const double* d_aabb1, const double* d_x2, int64_t n2,
                           const double* d_aabb2, int dim, const double* h_theta, const int64_t* d_indptr,
                           const int32_t* d_chunk, const double* d_noise_diag, int64_t row0, int32_t* d_indices,
                           double* d_data, void* stream) {
  WendlandParams p;
  int r = fill_wendland_params(p, d_x1, n1, d_aabb1, d_x2, n2, d_aabb2, dim, h_theta);
  if (r != 0) return r;
  if (n1 == 0) return 0;
  p.row0 = row0;
  p.indptr = (const long long*)d_indptr, p.noise = d_noise_diag, p.indices = d_indices, p.data = d_data;
  p.chunk = const_cast<int32_t*>(d_chunk);
  r = launch_wendland<true>(p, (cudaStream_t)stream);
  if (r != 0) return r;
  return launch_wendland_values(p, (cudaStream_t)stream);
}

int64_t fvgp_scan_scratch_len(int64_t n) { return (n + SCAN_CHUNK - 1) / SCAN_CHUNK + 2; }

int fvgp_exclusive_scan_i64(const int64_t* d_counts, int64_t n, int64_t* d_indptr, int64_t* d_scratch,
                            int64_t* h_total, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) {
    FVGP_CUDA_OK(cudaMemsetAsync(d_indptr, 0, sizeof(int64_t), st));
    if (h_total) *h_total = 0;
    return 0;
  }
  const long long nblk = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  launch(scan_block_sums_kernel, (unsigned)nblk, 256, 0, st, (const long long*)d_counts, n, (long long*)d_scratch);
  launch(scan_sums_kernel, 1, 32, 0, st, (long long*)d_scratch, nblk);
  launch(scan_apply_kernel, (unsigned)nblk, 256, 0, st, (const long long*)d_counts, n, (const long long*)d_scratch,
                                                    (long long*)d_indptr, nblk);
  FVGP_LAUNCH_OK();
  if (h_total) {
    FVGP_CUDA_OK(cudaMemcpyAsync(h_total, d_scratch + nblk, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    FVGP_CUDA_OK(cudaStreamSynchronize(st));
  }
  return 0;
}

int fvgp_csr_spmv(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                  const double* d_x, double* d_y, void* stream) {
  if (n <= 0) return 0;
  const long long want = (n * 8 + KR_THREADS - 1) / KR_THREADS;
  const unsigned grid = (unsigned)(want < (long long)krylov_grid() ? want : (long long)krylov_grid());
  FVGP_SPMV_DISPATCH(spmv_kernel, grid, KR_THREADS, 0, (cudaStream_t)stream, (long long)n, (const long long*)d_indptr,
                     d_indices, d_data, d_x, d_y);
  FVGP_LAUNCH_OK();
  return 0;
}

int64_t fvgp_bjacobi_len(int64_t n) { return ((n + 31) / 32) * 1024; }

int fvgp_bjacobi_build(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                       double* d_blocks, void* stream) {
  if (n <= 0) return 0;
  const long long nblk = (n + 31) / 32;
  launch(bjacobi_build_kernel, (unsigned)((nblk + 3) / 4), 128, 0, (cudaStream_t)stream, 
      n, (const long long*)d_indptr, d_indices, d_data, d_blocks);
  FVGP_LAUNCH_OK();
  return 0;
}

// work layout: r | z | p | q | partials (2*grid) | KrylovScalars
int64_t fvgp_pcg_work_len(int64_t n) { return 4 * n + 2 * (int64_t)krylov_grid() + 64; }

int fvgp_pcg(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
             const double* d_precond, const double* d_b, double* d_x, double rtol, int maxiter, double* d_work,
             int* h_iters, double* h_relres, void* stream) {
  FVGP_REQUIRE(n > 0 && maxiter > 0);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = krylov_grid();
  double* r = d_work;
  double* z = r + n;
  double* p = z + n;
  double* q = p + n;
  double* partials = q + n;
  KrylovScalars* sc = (KrylovScalars*)(partials + 2 * grid);
  const long long* ip = (const long long*)d_indptr;
  FVGP_CUDA_OK(cudaMemsetAsync(sc, 0, sizeof(KrylovScalars), st));
  FVGP_CUDA_OK(cudaMemsetAsync(p, 0, 2 * n * sizeof(double), st));  // p and q
  FVGP_SPMV_DISPATCH(pcg_init_kernel, grid, KR_THREADS, 0, st, (long long)n, ip, d_indices, d_data, d_b,
                     (const double*)d_x, r, rtol, sc, partials);
  FVGP_LAUNCH_OK();
  KrylovScalars h;
  const int batch = 16;
  // Three launches per iteration; the scalars (rho, alpha, beta, the convergence flag) never leave the
  // device, the host only polls the flag every `batch` iterations.  Kernels launched after convergence
  // return at once, so over-launching is harmless; pcg_head_kernel raises done = 2 after maxiter updates.
  for (long long launched = 0;; launched += batch) {
    FVGP_CUDA_OK(cudaMemcpyAsync(&h, sc, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, st));
    FVGP_CUDA_OK(cudaStreamSynchronize(st));
    if (h.done || launched > (long long)maxiter + batch) break;
    for (int k = 0; k < batch; ++k) {
      launch(pcg_head_kernel, grid, KR_THREADS, 0, st, n, d_precond, p, q, d_x, r, z, sc, partials, maxiter);
      launch(pcg_update_p_kernel, grid, KR_THREADS, 0, st, n, z, p, sc);
      FVGP_SPMV_DISPATCH(pcg_spmv_kernel, grid, KR_THREADS, 0, st, (long long)n, ip, d_indices, d_data,
                         (const double*)p, q, sc, partials);
    }
    FVGP_LAUNCH_OK();
  }
  if (h_iters) *h_iters = h.iters;
  if (h_relres) *h_relres = h.bnorm2 > 0.0 ? sqrt(h.rr / h.bnorm2) : 0.0;
  return h.done == 1 ? 0 : 1;
}

// work layout: p (n, full) | r | z | q (nrows each, padded to n) | partials (2*grid) | ShardScalars
int64_t fvgp_pcg_sharded_work_len(int64_t n) { return 4 * n + 2 * (int64_t)krylov_grid() + 64; }

int fvgp_pcg_sharded(void* comm, int64_t n, const int64_t* h_row_offsets, const int64_t* d_indptr_slab,
                     const int32_t* d_indices, const double* d_data, const double* d_precond_slab, const double* d_b,
                     double* d_x, double rtol, int maxiter, double* d_work, int* h_iters, double* h_relres,
                     void* stream) {
  FVGP_REQUIRE(comm != nullptr && n > 0 && maxiter > 0 && h_row_offsets != nullptr);
  const Comm* c = (const Comm*)comm;
  const long long row0 = h_row_offsets[c->rank], nrows = h_row_offsets[c->rank + 1] - row0;
  FVGP_REQUIRE(h_row_offsets[0] == 0 && h_row_offsets[c->world] == n && nrows >= 0 && row0 % 32 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = krylov_grid();
  double* p = d_work;  // full
  double* r = p + n;
  double* z = r + n;
  double* q = z + n;
  double* partials = q + n;
  ShardScalars* sc = (ShardScalars*)(partials + 2 * grid);
  const long long* ip = (const long long*)d_indptr_slab;
  int64_t off[65];
  FVGP_REQUIRE(c->world <= 64);
  for (int k = 0; k <= c->world; ++k) off[k] = h_row_offsets[k] * (int64_t)sizeof(double);
  FVGP_CUDA_OK(cudaMemsetAsync(sc, 0, sizeof(ShardScalars), st));
  FVGP_CUDA_OK(cudaMemsetAsync(p, 0, 4 * n * sizeof(double), st));
  FVGP_SPMV_DISPATCH(pcgs_init_kernel, grid, KR_THREADS, 0, st, nrows, ip, d_indices, d_data, d_b + row0,
                     (const double*)d_x, r, sc, partials);
  FVGP_LAUNCH_OK();
  int rc = comm_allreduce_sum(c, sc->loc, 2, st);
  if (rc != 0) return rc;
  launch(pcgs_scalar_kernel, 1, 32, 0, st, sc, 0, rtol, maxiter);
  ShardScalars h;
  const int batch = 16;
  for (long long launched = 0;; launched += batch) {
    FVGP_CUDA_OK(cudaMemcpyAsync(&h, sc, sizeof(ShardScalars), cudaMemcpyDeviceToHost, st));
    FVGP_CUDA_OK(cudaStreamSynchronize(st));
    if (h.done || launched > (long long)maxiter + batch) break;  // identical on every rank: same reduced scalars
    for (int k = 0; k < batch; ++k) {
      launch(pcgs_head_kernel, grid, KR_THREADS, 0, st, nrows, d_precond_slab, (const double*)(p + row0), (const double*)q,
             d_x + row0, r, z, sc, partials);
      if ((rc = comm_allreduce_sum(c, sc->loc, 2, st)) != 0) return rc;
      launch(pcgs_scalar_kernel, 1, 32, 0, st, sc, 1, rtol, maxiter);
      launch(pcgs_update_p_kernel, grid, KR_THREADS, 0, st, nrows, (const double*)z, p + row0, (const ShardScalars*)sc);
      if ((rc = comm_allgatherv_bytes(c, p, off, st)) != 0) return rc;
      FVGP_SPMV_DISPATCH(pcgs_spmv_kernel, grid, KR_THREADS, 0, st, nrows, ip, d_indices, d_data, (const double*)p,
                         (const double*)(p + row0), q, sc, partials);
      if ((rc = comm_allreduce_sum(c, sc->loc + 2, 1, st)) != 0) return rc;
      launch(pcgs_scalar_kernel, 1, 32, 0, st, sc, 2, rtol, maxiter);
    }
    FVGP_LAUNCH_OK();
  }
  if ((rc = comm_allgatherv_bytes(c, d_x, off, st)) != 0) return rc;  // every rank leaves with the whole solution
  FVGP_CUDA_OK(cudaStreamSynchronize(st));
  if (h_iters) *h_iters = h.iters;
  if (h_relres) *h_relres = h.bnorm2 > 0.0 ? sqrt(h.rr / h.bnorm2) : 0.0;
  return h.done == 1 ? 0 : 1;
}

// work layout: three n x 16 blocks (u_prev | u | w, rotating) | partials (16*grid) | LanczosScalars | alpha, beta
int64_t fvgp_lanczos_work_len(int64_t n, int degree) {
  return 3 * n * LZ_MAXB + (int64_t)LZ_MAXB * krylov_grid() + 128 + 2 * (int64_t)degree * LZ_MAXB;
}

}  // extern "C"

// FVGP_LANCZOS_COLS=0 selects the per-lane-entry SpMM for batches of >= 4 probes (A/B on the GPU box).
static int lanczos_cols_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FVGP_LANCZOS_COLS");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v;
}

template <int NB>
static int lanczos_batch(int64_t n, const long long* ip, const int32_t* d_indices, const double* d_data, int degree,
                         int probe0, uint64_t seed, double* d_work, double* h_alpha, double* h_beta, cudaStream_t st,
                         int keep = NB) {  // keep < NB: a padded batch, only the first `keep` probes are returned
  const unsigned grid = krylov_grid();
  double* bufs[3] = {d_work, d_work + n * LZ_MAXB, d_work + 2 * n * LZ_MAXB};
  double* partials = d_work + 3 * n * LZ_MAXB;
  LanczosScalars* sc = (LanczosScalars*)(partials + (size_t)LZ_MAXB * grid);
  double* d_alpha = (double*)(sc) + 128 - 0;  // past the scalars block (sizeof(LanczosScalars) <= 128 doubles)
  double* d_beta = d_alpha + (size_t)degree * LZ_MAXB;
  static_assert(sizeof(LanczosScalars) <= 128 * sizeof(double), "scalars block");
  FVGP_CUDA_OK(cudaMemsetAsync(sc, 0, sizeof(LanczosScalars), st));
  double *uprev = bufs[0], *u = bufs[1], *w = bufs[2];
  launch(lanczos_start_kernel<NB>, grid, KR_THREADS, 0, st, (long long)n, (unsigned long long)seed,
         (unsigned long long)probe0, u, uprev, sc);
  for (int j = 0; j < degree; ++j) {
    if constexpr (NB >= 4) {
      if (lanczos_cols_variant())
        launch(lanczos_spmm_cols_kernel<NB>, grid, KR_THREADS, 0, st, (long long)n, ip, d_indices, d_data, u, uprev, w, sc,
               partials);
      else
        launch(lanczos_spmm_kernel<NB>, grid, KR_THREADS, 0, st, (long long)n, ip, d_indices, d_data, u, uprev, w, sc,
               partials);
    } else {
      launch(lanczos_spmm_kernel<NB>, grid, KR_THREADS, 0, st, (long long)n, ip, d_indices, d_data, u, uprev, w, sc,
             partials);
    }
    launch(lanczos_axpy_kernel<NB>, grid, KR_THREADS, 0, st, (long long)n, u, w, sc, partials, d_alpha + (size_t)j * NB,
           d_beta + (size_t)j * NB);
    double* t = uprev;
    uprev = u, u = w, w = t;
  }
  FVGP_LAUNCH_OK();
  // device layout [j][c] -> host layout [probe][j]
  static thread_local double ha[64 * LZ_MAXB], hb[64 * LZ_MAXB];
  FVGP_REQUIRE(degree <= 64);
  FVGP_CUDA_OK(cudaMemcpyAsync(ha, d_alpha, (size_t)degree * NB * sizeof(double), cudaMemcpyDeviceToHost, st));
  FVGP_CUDA_OK(cudaMemcpyAsync(hb, d_beta, (size_t)degree * NB * sizeof(double), cudaMemcpyDeviceToHost, st));
  FVGP_CUDA_OK(cudaStreamSynchronize(st));
  for (int c = 0; c < keep; ++c)
    for (int j = 0; j < degree; ++j) {
      h_alpha[(size_t)c * degree + j] = ha[(size_t)j * NB + c];
      h_beta[(size_t)c * degree + j] = hb[(size_t)j * NB + c];
    }
  return 0;
}

extern "C" {

int fvgp_lanczos_tridiag(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                         int degree, int probe0, int nprobes, uint64_t seed, double* d_work, double* h_alpha,
                         double* h_beta, void* stream) {
  FVGP_REQUIRE(n > 0 && degree > 0 && degree <= 64 && nprobes >= 0);
  cudaStream_t st = (cudaStream_t)stream;
  const long long* ip = (const long long*)d_indptr;
  int done = 0;
  // Batch policy, from the B200 measurements at C4 (profiles/r02): one sweep over the matrix per Lanczos step costs
  // about the same for 1 ... 8 probe columns (~0.7 ms: bound by streaming the matrix and by latency, not by the column
  // count -- 8 + 2 probes took the time of 4 + 1), while the 16-column kernel is slower than two 8-column sweeps
  // (1.8 vs 1.4 ms).  So: batches of at most 8, and a remainder runs as ONE batch padded to the next power of two
  // (the surplus columns are later probes of the same stream that are not returned) instead of 4 + 2 + 1.
  // FVGP_SLQ_GREEDY=1 restores the greedy 16 / 8 / 4 / 2 / 1 split (A/B).
  static int greedy = -1;
  if (greedy < 0) {
    const char* e = getenv("FVGP_SLQ_GREEDY");
    greedy = (e && atoi(e) == 1) ? 1 : 0;
  }
  while (done < nprobes) {
    const int left = nprobes - done;
    int nb, keep;
    if (greedy) {
      nb = left >= 16 ? 16 : left >= 8 ? 8 : left >= 4 ? 4 : left >= 2 ? 2 : 1;
      keep = nb;
    } else {
      nb = 1;
      while (nb < left && nb < 8) nb *= 2;
      keep = left < nb ? left : nb;
    }
    double* ha = h_alpha + (size_t)done * degree;
    double* hb = h_beta + (size_t)done * degree;
    int r;
    switch (nb) {
      case 16: r = lanczos_batch<16>(n, ip, d_indices, d_data, degree, probe0 + done, seed, d_work, ha, hb, st, keep); break;
      case 8: r = lanczos_batch<8>(n, ip, d_indices, d_data, degree, probe0 + done, seed, d_work, ha, hb, st, keep); break;
      case 4: r = lanczos_batch<4>(n, ip, d_indices, d_data, degree, probe0 + done, seed, d_work, ha, hb, st, keep); break;
      case 2: r = lanczos_batch<2>(n, ip, d_indices, d_data, degree, probe0 + done, seed, d_work, ha, hb, st, keep); break;
      default: r = lanczos_batch<1>(n, ip, d_indices, d_data, degree, probe0 + done, seed, d_work, ha, hb, st, keep); break;
    }
    if (r != 0) return r;
    done += keep;
  }
  return 0;
}

}  // extern "C"
