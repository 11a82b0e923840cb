/* fvgp_b200 -- C ABI of the B200-native fvGP training hot path.
 *
 * Drop-in boundary.  Every entry point replaces one Python-level
 * function of the reference on the LML(+gradient) path; the reference function is cited
 * as file:line into lbl-camera/fvGP.  The reference is pure Python, so the binding a
 * maintainer adds is a ctypes stub; `fvgp_b200/_lib.py` is ours.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types.
 *  - pointers named d_* are DEVICE pointers (owned by the caller, e.g. torch tensors);
 *    pointers named h_* are HOST pointers (small parameter vectors / scalar results).
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it.  Functions
 *    that return a scalar through an h_* pointer synchronise that stream once.
 *  - matrices are C-order (row-major) float64 with leading dimension ld* in elements;
 *    ld must be even and bases 16-byte aligned.
 *  - return value: 0 ok; >0 = 1-based index of the first non-positive pivot (-> the
 *    reference's NonPositiveDefiniteError, gp_lin_alg.py:27-58); <0 CUDA / argument error.
 *  - no global mutable state besides one-time kernel attribute set-up; re-entrant across
 *    streams of one device.
 */
#ifndef FVGP_B200_H
#define FVGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- radial kernel families: value = amp * f(d / length), d = ||(x1 - x2) * inv_scale|| */
enum fvgp_kernel_kind {
  FVGP_K_MATERN32 = 0, /* kernels.py:98-118  matern_kernel_diff1; default kernel gp_prior.py:376-400 */
  FVGP_K_MATERN52 = 1, /* kernels.py:166-188 matern_kernel_diff2 */
  FVGP_K_SQEXP = 2,    /* kernels.py:16-33   squared_exponential_kernel */
  FVGP_K_EXP = 3,      /* kernels.py:56-74   exponential_kernel */
  FVGP_K_WENDLAND = 4, /* kernels.py:355-378 wendland_anisotropic (dense form) */
  FVGP_K_DISTANCE = 5, /* kernels.py:440-481 get_(anisotropic_)distance_matrix: value = d */
  FVGP_K_MATERN52_ROBUST = 6 /* kernels.py:191-213 matern_kernel_diff2_robust with length = 1 / phi^2: the reference's
                              * quadratic term is (5 d^2)(3 phi^4) = 15 (d / length)^2, not 5/3 -- reproduced as is */
};

enum fvgp_fill_mode {
  FVGP_FILL_FULL = 0,      /* every entry of the n1 x n2 matrix */
  FVGP_FILL_SYMMETRIC = 1, /* x1 == x2: evaluate upper tiles once, write tile + transpose */
  FVGP_FILL_LOWER = 2      /* x1 == x2: tiles on/below the diagonal only (input of potrf) */
};

int fvgp_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long fvgp_launch_count(void);
/* test / profiling hook: mirror tiles of the symmetric K-fill through TMA bulk stores (1, default)
 * or plain coalesced stores (0).  Returns the previous setting. */
int fvgp_set_bulk_store(int on);

/* Dense covariance assembly fused with the noise diagonal.
 * Replaces GPprior._default_kernel (gp_prior.py:376-400), kernels.get_distance_matrix /
 * get_anisotropic_distance_matrix (kernels.py:440-481) + the radial kernels
 * (kernels.py:16-188), and GPkv.addKV for vector noise (gp_kv.py:640-669).
 * h_inv_scale[dim]: 1/length-scale per axis (all 1.0 = isotropic).  d_noise may be NULL.
 * h_centre[dim] (may be NULL): a point c such that |x - c| * inv_scale / length * sqrt(5) <= 512 for every
 * point of x1 and x2.  When given (and dim <= 4) the coordinates are centred and scaled once per tile and
 * the per-entry distance is a difference of scaled coordinates (3 fewer FP64 operations per axis; the
 * entry's extra relative error is bounded by ~2e-13).  NULL selects the reference's operation order
 * ((x1 - x2) * inv_scale per entry). */
int fvgp_kfill_dense(int kind, int mode, const double* d_x1, int64_t n1, const double* d_x2, int64_t n2, int dim,
                     double amp, const double* h_inv_scale, const double* h_centre, double length,
                     const double* d_noise, double* d_K, int64_t ldk, void* stream);

/* The radial kernels of kernels.py:16-188 applied elementwise to a caller-supplied distance array. */
int fvgp_radial_elementwise(int kind, const double* d_dist, int64_t count, double amp, double length, double* d_out,
                            void* stream);

/* Gradient traces without materialising dK/dtheta (gp_marginal_likelihood.py:256-309 with
 * gp_prior.py:421-436): for the default ARD Matern-3/2 kernel and theta = (amp, l_1..l_dim)
 *   h_out[h] = sum_ij W_ij * dK_ij/dtheta_h,   W = Kinv - b b^T  (lower triangle of d_Kinv is read)
 * so that grad(-LML)_h = 0.5 * h_out[h].  d_partials: workspace of fvgp_kgrad_partials_len() doubles. */
int64_t fvgp_kgrad_partials_len(int64_t n, int dim);
int fvgp_kgrad_trace_matern32(const double* d_x, int64_t n, int dim, const double* h_theta, const double* d_Kinv,
                              int64_t ld, const double* d_b, double* d_partials, double* h_out, void* stream);

/* The same traces for every radial family with a fused form (Matern-3/2, Matern-5/2, squared exponential,
 * exponential; kernels.py:16-188) and K = amp * f(||(x1 - x2) * inv_scale|| / length), with respect to the
 * descriptor's own parameters:
 *   h_out[0]       = sum_ij W_ij dK_ij/d(amp)
 *   h_out[1 + i]   = sum_ij W_ij dK_ij/d(inv_scale_i)     (i < dim)
 *   h_out[dim + 1] = sum_ij W_ij dK_ij/d(length)
 * A user kernel composed from the fvgp.kernels names maps its hyperparameters to (amp, inv_scale, length); the
 * host applies the chain rule (replaces the finite-difference dK of gp_prior.py:438-447, which materialises
 * H dense N x N arrays).  d_partials: fvgp_kgrad_partials_len(n, dim) doubles. */
int fvgp_kgrad_trace_radial(int kind, const double* d_x, int64_t n, int dim, double amp, const double* h_inv_scale,
                            double length, const double* d_Kinv, int64_t ld, const double* d_b, double* d_partials,
                            double* h_out, void* stream);

/* One m x n block of the same trace for the block-cyclic multi-GPU layout: rows belong to (x1, b1), columns to
 * (x2, b2); the first diag_rows rows are a diagonal block aligned with the columns (lower triangle only, diagonal
 * weighted once), all other entries are weighted twice.  d_accum[h] (dim+1 doubles, device) += the block's
 * contribution; no host synchronisation.  d_partials: fvgp_kgrad_block_partials_len(dim) doubles.
 * NOTE: the memcpy of 1/length from a stack buffer is stream-ordered with pageable memory (synchronous copy-in). */
int64_t fvgp_kgrad_block_partials_len(int dim);
int fvgp_kgrad_trace_block_matern32(const double* d_x1, int64_t m, const double* d_x2, int64_t n, int dim,
                                    const double* h_theta, const double* d_W, int64_t ldw, const double* d_b1,
                                    const double* d_b2, int64_t diag_rows, double* d_partials, double* d_accum,
                                    void* stream);

/* The same block trace for every radial family with fused gradient traces (kinds MATERN32, MATERN52, SQEXP, EXP),
 * i.e. for user kernels composed of the fvgp.kernels names (kernels.py:16-188) on the sharded dense path.
 * d_accum_raw (dim + 1 doubles, device) is INCREMENTED by the raw sums R_0 = sum W f, R_i = sum W h q_i; after the
 * all-reduce the traces against (amp, inv_scale_1..D, length) are T_amp = R_0, T_si = -(amp / s_i) R_i,
 * T_length = (amp / length) sum_i R_i -- the conversion fvgp_kgrad_trace_radial applies on one GPU. */
int fvgp_kgrad_trace_block_radial(int kind, const double* d_x1, int64_t m, const double* d_x2, int64_t n, int dim,
                                  const double* h_inv_scale, double length, const double* d_W, int64_t ldw,
                                  const double* d_b1, const double* d_b2, int64_t diag_rows, double* d_partials,
                                  double* d_accum_raw, void* stream);

/* Same trace against a MATERIALISED symmetric dK (user kernel_function_grad, gp_prior.py:236-240):
 * *h_out = sum_ij (Kinv - b b^T)_ij dK_ij.  d_partials: 148*8 + 1 doubles. */
int fvgp_trace_sym_product(const double* d_Kinv, int64_t ld, const double* d_b, const double* d_dK, int64_t lddk,
                           int64_t n, double* d_partials, double* h_out, void* stream);

/* Dense dK/dtheta materialised, (dim+1) x n1 x n2 (GPprior._default_kernel_analytical_gradient,
 * gp_prior.py:421-436) -- the kernel_function_grad seam (gp_prior.py:236-240). */
int fvgp_kgrad_dense_matern32(const double* d_x1, int64_t n1, const double* d_x2, int64_t n2, int dim,
                              const double* h_theta, double* d_out, void* stream);

/* ---- dense FP64 factorisation (gp_lin_alg.py:237-360, :1558) */
int64_t fvgp_chol_workspace_len(int64_t n);  /* doubles: per-128-tile inverses kept by potrf  */
int64_t fvgp_potri_workspace_len(int64_t n); /* doubles: scratch panel of trtri / lauum       */

/* calculate_Chol_factor (gp_lin_alg.py:237-269): in-place lower Cholesky of the lower
 * triangle of d_A.  d_tileinv (fvgp_chol_workspace_len doubles) receives the inverses of the
 * 128x128 diagonal tiles (tile t at offset t*128*128) and must be passed unchanged to potrs / potri.  d_info: one int. */
int fvgp_potrf_lower(double* d_A, int64_t n, int64_t lda, double* d_tileinv, int* d_info, void* stream);
/* The same factorisation without the host synchronisation: the status stays in *d_info (0, or the 1-based failing
 * pivot) for the caller to read later -- the block-cyclic factorisation reads all its panels' flags once at the end. */
int fvgp_potrf_lower_enqueue(double* d_A, int64_t n, int64_t lda, double* d_tileinv, int* d_info, void* stream);

/* calculate_Chol_solve (gp_lin_alg.py:289-328): solve (L L^T) X = B in place.  d_B holds
 * nrhs right-hand sides, each contiguous with stride ldb (i.e. B^T in C order).
 * d_work: fvgp_potrs_work_len(n) doubles.  Up to 4 right-hand sides run as blocked substitutions whose
 * off-diagonal panels are matrix-vector products (HBM-read bound: the factor is read once per direction);
 * more go through the tensor-core TRSM recursion. */
int64_t fvgp_potrs_work_len(int64_t n);
int fvgp_potrs_lower(const double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_B, int nrhs,
                     int64_t ldb, double* d_work, void* stream);

/* calculate_Chol_logdet (gp_lin_alg.py:331-360): 2 * sum log|L_ii| -> *h_out. */
int fvgp_chol_logdet(const double* d_L, int64_t n, int64_t lda, double* d_scratch1, double* h_out, void* stream);

/* calculate_inv_from_chol (gp_lin_alg.py:1558) / the trace term's KV^-1
 * (gp_marginal_likelihood.py:273-274): lower triangle of d_L <- lower triangle of (L L^T)^-1. */
int fvgp_potri_lower(double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_work, void* stream);

/* ---- population evaluation: B hyperparameter proposals of one radial kernel family on the same points.
 * Replaces the per-individual objective calls of the optimisers -- scipy differential_evolution generations
 * (gp_training.py:60-80), hgdl / multi-start local walkers, the H+1 gradient calls of the finite-difference Hessian
 * (gp_marginal_likelihood.py:312-336) -- each of which is one compute_new_KVlogdet_KVinvY (gp_kv.py:574-631) plus,
 * with want_grad, one neg_log_likelihood_gradient (gp_marginal_likelihood.py:224-309).  At the sizes where training
 * loops are hot (N ~ 1e3..1e4) one evaluation is a latency-bound chain of small launches that leaves the GPU nearly
 * empty; here proposal b runs K-fill -> POTRF -> POTRS -> logdet (-> POTRI -> fused traces) on stream b % slots with
 * its own workspace slot, so up to `slots` chains overlap, and the host synchronises ONCE for the whole population.
 * The arithmetic of every proposal is exactly that of the one-at-a-time entry points (same kernels, same order).
 *   kind, h_amp[b], h_inv_scale[b*dim + i], h_length[b]: K_b = amp_b * f(||(x - x') * inv_scale_b|| / length_b)
 *   h_centre: dim doubles or NULL, as fvgp_kfill_dense (must be valid for every proposal)
 *   d_noise:  n doubles or NULL, added to the diagonal;  d_rhs: nrhs x n (row = one right-hand side, y - m), nrhs <= 4
 *   d_work:   slots * fvgp_population_slot_len(n, dim, want_grad) doubles
 *   d_alpha:  batch * nrhs * n doubles (device), h_alpha the same on the host: KV_b^-1 rhs
 *   d_res:    batch * (dim + 2) doubles, d_info: batch ints (device scratch)
 *   h_logdet[b] = log|KV_b|;  h_info[b] = 0, or the 1-based failing pivot when KV_b is not positive definite (its other
 *   outputs are then undefined);  h_traces[b*(dim+2) ..] (want_grad) = sums of (KV_b^-1 - a a^T) o dK/dp over
 *   p = (amp, inv_scale_1..dim, length) with a = column grad_component of alpha_b, as fvgp_kgrad_trace_radial. */
int64_t fvgp_population_slot_len(int64_t n, int dim, int want_grad);
int fvgp_lml_population(int kind, const double* d_x, int64_t n, int dim, int batch, const double* h_amp,
                        const double* h_inv_scale, const double* h_length, const double* h_centre,
                        const double* d_noise, const double* d_rhs, int nrhs, int want_grad, int grad_component,
                        int slots, double* d_work, double* d_alpha, double* d_res, int* d_info, double* h_alpha,
                        double* h_logdet, double* h_traces, int* h_info, void* stream);

/* ---- building blocks of the 2-D block-cyclic multi-GPU factorisation (fvgp_b200/sharded.py).
 * Each is the single-GPU piece of one step of the distributed POTRF / POTRS / POTRI that replaces
 * calculate_Chol_factor / calculate_Chol_solve / the gradient's KV^-1 when KV exceeds one HBM. */
enum fvgp_gemm_flags {
  FVGP_GEMM_LOWER = 1,     /* square tile grid; tiles strictly above the diagonal are skipped */
  FVGP_GEMM_KB_FROM_M = 2, /* k starts at the tile's first row    (A^T lower triangular) */
  FVGP_GEMM_KB_FROM_N = 4, /* k starts at the tile's first column (B lower triangular)   */
  FVGP_GEMM_KE_FROM_M = 8  /* k ends at the tile's last row       (A lower triangular)   */
};
/* C (m x n, row-major) = alpha * op(A) * op(B) + beta * C on the FP64 tensor cores (DMMA).
 * a_mn = 0: A is m x k row-major (A[m][k]); a_mn = 1: A is stored k x m (A[k][m]).
 * b_mn = 0: B is n x k row-major (B[n][k], i.e. C += A B^T); b_mn = 1: B is stored k x n.
 * (a_mn, b_mn) = (1, 0) is not instantiated. */
int fvgp_dgemm(int a_mn, int b_mn, const double* d_A, int64_t lda, const double* d_B, int64_t ldb, double* d_C,
               int64_t ldc, int m, int n, int k, double alpha, double beta, int flags, void* stream);
/* B (m x n) <- B * L^-T with L the n x n lower factor produced by fvgp_potrf_lower (+ its tile inverses). */
int fvgp_trsm_right_lower_t(double* d_B, int64_t ldb, int m, const double* d_L, int64_t ldl, int n,
                            const double* d_tileinv, void* stream);
/* The two halves of fvgp_potri_lower: L <- L^-1 (lower), and lower(M) <- lower(M^T M). */
int fvgp_trtri_lower(double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_work, void* stream);
int fvgp_lauum_lower(double* d_M, int64_t n, int64_t lda, double* d_work, void* stream);
/* One triangular solve against the factor, in place on d_b: transpose = 0: L z = b; 1: L^T x = b.
 * d_work: 2*n doubles. */
int fvgp_trsv_lower(const double* d_L, int64_t n, int64_t lda, const double* d_tileinv, double* d_b, int transpose,
                    double* d_work, void* stream);
/* y += alpha * A x (transpose = 0, A m x n row-major, x[n], y[m]) or y += alpha * A^T x (transpose = 1,
 * x[m], y[n]); deterministic.  d_work: fvgp_gemv_work_len(m, n) doubles (transpose = 1 only). */
int64_t fvgp_gemv_work_len(int64_t m, int64_t n);
int fvgp_gemv(int transpose, const double* d_A, int64_t lda, int m, int n, double alpha, const double* d_x,
              double* d_y, double* d_work, void* stream);

/* Plain tensor-core GEMM exposed for tests / roofline measurement:
 * C = alpha * A * B^T + beta * C with A (m x k) and B (n x k) row-major. lower!=0: lower tiles only. */
int fvgp_dgemm_nt(const double* d_A, int64_t lda, const double* d_B, int64_t ldb, double* d_C, int64_t ldc, int m,
                  int n, int k, double alpha, double beta, int lower, void* stream);

/* sum_i a_i * b_i -> *h_out (quadratic form of gp_marginal_likelihood.py:175). */
int fvgp_dot(const double* d_a, const double* d_b, int64_t n, double* d_scratch1, double* h_out, void* stream);

/* ---- gp2Scale: compact-support covariance straight to canonical CSR
 * Replaces wendland_anisotropic_gp2Scale_cpu (kernels.py:502-528) evaluated per block by
 * block_triplets / block_to_coo (gp2Scale_covariance.py:136-170) and the host assembly
 * assemble_triplets (gp2Scale_covariance.py:240-287), and addKV's setdiag (gp_kv.py:655-661).
 * Pass 1 counts entries per row; the caller scans counts into d_indptr (int64 or int32 is
 * the caller's choice for the final matrix; the kernels use int64 offsets); pass 2 fills
 * sorted column indices (same geometry pass) and then the values (lane-dense kernel over the stored
 * entries, reference operation order).  Pattern is bit-exact w.r.t. the reference predicate.
 * Work is cut into (32-row tile, column chunk) units, at most 32 chunks; d_chunk (fvgp_wendland_chunk_len
 * int32) carries the per-(chunk, row) counts from the count pass to the fill pass and must not be touched
 * in between.
 * d_stats (may be NULL): TWO int64 the count pass ADDS to: [0] the number of (row tile, column tile) pairs whose
 * boxes survived both culls, [1] the number of candidate point pairs actually tested (a per-row cull against the
 * column tile's box runs in front of the pair loop) -- reported next to nnz as the cull efficiency.
 * Row slabs (multi-GPU row sharding; gp2Scale_covariance.py:381-396 rowwise tasks): x1 may be rows
 * [row0, row0 + n1) of x2; d_indptr then points at the slab's n1 + 1 offsets (absolute positions inside d_indices /
 * d_data), d_noise_diag at the slab's n1 noise values, and row0 tells the fill where the diagonal is. */
int64_t fvgp_wendland_aabb_len(int64_t n, int dim); /* doubles per point set for tile bounding boxes */
int fvgp_wendland_aabb(const double* d_x, int64_t n, int dim, double* d_aabb, void* stream);
int64_t fvgp_wendland_chunk_len(int64_t n1, int64_t n2); /* int32 scratch shared by count and fill */
int fvgp_wendland_csr_count(const double* d_x1, int64_t n1, const double* d_aabb1, const double* d_x2, int64_t n2,
                            const double* d_aabb2, int dim, const double* h_theta, int64_t* d_rowcount,
                            int32_t* d_chunk, int64_t* d_stats, void* stream);
int fvgp_wendland_csr_fill(const double* d_x1, int64_t n1, const double* d_aabb1, const double* d_x2, int64_t n2,
                           const double* d_aabb2, int dim, const double* h_theta, const int64_t* d_indptr,
                           const int32_t* d_chunk, const double* d_noise_diag, int64_t row0, int32_t* d_indices,
                           double* d_data, void* stream);
/* exclusive scan of n counts into n+1 offsets (d_indptr[0] = 0); *h_total = nnz. */
int fvgp_exclusive_scan_i64(const int64_t* d_counts, int64_t n, int64_t* d_indptr, int64_t* d_scratch,
                            int64_t* h_total, void* stream);
int64_t fvgp_scan_scratch_len(int64_t n);

/* ---- sparse solve (gp_lin_alg.py:1213-1291 calculate_sparse_conj_grad -> scipy cg) */
int fvgp_csr_spmv(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                  const double* d_x, double* d_y, void* stream);

/* Block-Jacobi preconditioner: inverses of the dense bs x bs diagonal blocks of the CSR
 * matrix (cf. calculate_sparse_preconditioner "block_jacobi", gp_lin_alg.py:604-622). bs = 32. */
int64_t fvgp_bjacobi_len(int64_t n);
int fvgp_bjacobi_build(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                       double* d_blocks, void* stream);

/* Preconditioned CG, scipy semantics: x0 given in d_x, stop when ||r||_2 < rtol*||b||_2
 * (checked at the top of each iteration), atol = 0.  d_precond: fvgp_bjacobi_build output
 * or NULL (plain CG).  d_work: fvgp_pcg_work_len(n) doubles.
 * h_iters / h_relres receive iteration count and final relative residual.
 * Returns 0 on convergence, 1 if maxiter was reached (scipy's info > 0). */
int64_t fvgp_pcg_work_len(int64_t n);
int fvgp_pcg(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
             const double* d_precond, const double* d_b, double* d_x, double rtol, int maxiter, double* d_work,
             int* h_iters, double* h_relres, void* stream);

/* Stochastic Lanczos quadrature log-determinant (calculate_random_logdet,
 * gp_lin_alg.py:1103-1181 -> imate slq): Rademacher probes (counter-based, `seed`),
 * `degree` Lanczos steps each, full reorthogonalisation off (imate orthogonalize=0).
 * Device part: the Lanczos recurrences; outputs the tridiagonal coefficients
 * h_alpha[probe*degree + j], h_beta[probe*degree + j] for the host-side quadrature.
 * d_work: fvgp_lanczos_work_len(n, degree) doubles. */
int64_t fvgp_lanczos_work_len(int64_t n, int degree);
int fvgp_lanczos_tridiag(int64_t n, const int64_t* d_indptr, const int32_t* d_indices, const double* d_data,
                         int degree, int probe0, int nprobes, uint64_t seed, double* d_work, double* h_alpha,
                         double* h_beta, void* stream);

/* ---- multi-GPU (SURVEY 8b "multi-GPU variants take an ncclComm_t", 8e).  One process per GPU; the reference's
 * counterpart is the dask cluster of gp2Scale (gp2Scale_covariance.py:313-431: blockwise / rowwise tasks, scatter of
 * x, gp_prior.py:301-322) -- it has no multi-GPU solver.  The library does not link NCCL: fvgp_nccl_attach binds the
 * NCCL the calling process already uses (path of the loaded libnccl.so.2, or NULL for the SONAME).
 * A comm handle is either created here (rank 0 draws fvgp_comm_unique_id, the host distributes the 128 bytes -- e.g.
 * through torch.distributed -- and every rank calls fvgp_comm_create) or adopts an ncclComm_t the caller owns.
 * All collective entry points must be called by every rank of the communicator, in the same order. */
int fvgp_nccl_attach(const char* libnccl_path);
int fvgp_comm_unique_id(void* h_id128);
int fvgp_comm_create(const void* h_id128, int rank, int world, void** comm_out);
int fvgp_comm_adopt(void* nccl_comm, int rank, int world, void** comm_out);
int fvgp_comm_destroy(void* comm);
/* In-place all-gather of unequal parts: rank r owns bytes [h_offsets_bytes[r], h_offsets_bytes[r+1]) of d_buf
 * (row slabs of the gp2Scale CSR: assemble_row_strips, gp2Scale_covariance.py:290-296).  One grouped NCCL launch. */
int fvgp_comm_allgatherv(void* comm, void* d_buf, const int64_t* h_offsets_bytes, void* stream);
int fvgp_comm_allreduce_sum(void* comm, double* d_buf, int64_t count, void* stream);

/* Row-sharded preconditioned CG (calculate_sparse_conj_grad, gp_lin_alg.py:1213-1291, same stopping rule as
 * fvgp_pcg): rank r owns rows [h_row_offsets[r], h_row_offsets[r+1]) (multiples of 32) of the matrix, of the
 * block-Jacobi blocks and of x / r / z / q; the search direction lives in full on every rank.  d_indptr_slab: the
 * slab's nrows + 1 offsets (absolute positions in d_indices / d_data), d_precond_slab: the slab's 32 x 32 blocks or
 * NULL, d_b: full right-hand side, d_x: full vector (x0 in, the whole solution out on every rank).  Per iteration:
 * slab SpMV + fused vector kernels + two scalar all-reduces + one all-gather of p, all enqueued on `stream`; the host
 * polls the convergence flag every 16 iterations.  d_work: fvgp_pcg_sharded_work_len(n) doubles. */
int64_t fvgp_pcg_sharded_work_len(int64_t n);
int fvgp_pcg_sharded(void* comm, int64_t n, const int64_t* h_row_offsets, const int64_t* d_indptr_slab,
                     const int32_t* d_indices, const double* d_data, const double* d_precond_slab, const double* d_b,
                     double* d_x, double rtol, int maxiter, double* d_work, int* h_iters, double* h_relres,
                     void* stream);

/* ---- FP64-accurate GEMM on the INT8 tensor cores (tcgen05.mma kind::i8 + TMEM; csrc/ozaki.cu): the trailing updates
 * of calculate_Chol_factor (gp_lin_alg.py:237-269 -> LAPACK dpotrf) past the DMMA roof.  Ozaki-type splitting: rows
 * scaled by powers of two, `slices` 6-bit digits per entry, exact int32 slice products grouped by scale into `slices`
 * int8 GEMMs, FP64 recombination.  slices = 8: ~2^-47 relative to the row maxima, 36 integer MACs per FP64 MAC.
 * C (m x n, ldc) += sign * A (m x k, lda) B (n x k, ldb)^T, FP64 row-major; lower != 0 updates only the entries with
 * column <= row + diag; same_ab != 0: B is A (SYRK).  k % 16 == 0.  C is processed in column blocks of nblock.
 * fvgp_ozaki_available() = 0 when the library was built without the CuTe / CUTLASS headers.
 * Returns 0; < 0 before any entry of C was touched (arguments / a GEMM the hardware path refuses: C is intact and the
 * caller may fall back to fvgp_dgemm); -100 if a GEMM failed after part of C had been updated. */
int fvgp_ozaki_available(void);
/* 0: DMMA trailing updates in fvgp_potrf_lower; 6..10: INT8-slice updates with that many slices for updates of at
 * least 8192 rows in factorisations with 2048-wide block columns (N >= 40 000).  Default 8 when the library was built
 * with the CuTe / CUTLASS headers (FVGP_OZAKI=0 in the environment switches it off).  Returns the previous setting. */
int fvgp_set_ozaki(int slices);
/* 0: inside fvgp_potri_lower only the SYRK half of LAUUM uses the INT8-slice path; c > 0 (default 8): also the products
 * with a triangular operand (both TRTRI products and W = M22^T M21 of LAUUM; the inverse is the reference's
 * calculate_inv_from_chol / np.linalg.inv, gp_lin_alg.py:1540-1558, used by the trace term of the gradient at
 * gp_marginal_likelihood.py:273-274), with the contraction range cut into c chunks so that an int8 GEMM only multiplies
 * the part the triangle reaches.  FVGP_OZAKI_TRI in the environment sets the start value.  Returns the previous setting. */
int fvgp_set_ozaki_tri(int chunks);
/* Which factorisations / inversions use the INT8-slice products: N >= min_n (default 40 000, the range it was measured
 * to pay in; below it fvgp_potrf_lower also needs an update of >= 8192 rows) and, inside fvgp_potri_lower, per recursion
 * level a leading block of at least min_rows rows (default 8192, >= 256).  Lowered by the parity tests so that the path
 * is compared with the oracle at sizes the oracle finishes in seconds. */
int fvgp_set_ozaki_gate(int min_n, int min_rows);
/* The INT8-slice product with fvgp_dgemm's operand layouts (a_mn / b_mn as there; (1, 0) is accepted here), for the
 * trailing updates of the block-cyclic multi-GPU factorisation and inversion (fvgp_b200/sharded.py):
 * C (m x n) = sign * op(A) op(B) + (zero_c ? 0 : C).  Operands stored k x rows are transposed into scratch (k padded
 * to 16); a K-major operand needs k % 16 == 0.  Return values as fvgp_ozaki_gemm_nt (after a refusal with zero_c the
 * caller redoes the product with fvgp_dgemm and beta = 0).  fvgp_ozaki_slices(): the current fvgp_set_ozaki setting. */
int fvgp_ozaki_slices(void);
int64_t fvgp_ozaki_gemm_work_bytes(int a_mn, int b_mn, int64_t m, int64_t n, int64_t k, int slices, int64_t nblock);
int fvgp_ozaki_gemm(int a_mn, int b_mn, const double* d_A, int64_t lda, const double* d_B, int64_t ldb, double* d_C,
                    int64_t ldc, int64_t m, int64_t n, int64_t k, double sign, int zero_c, int slices, int64_t nblock,
                    void* d_work, int64_t work_bytes, void* stream);
/* Measurement hook: seconds per launch of the raw int8 GEMM (m x n x K, int32 output) behind fvgp_ozaki_gemm_nt for
 * tile configuration 1..4 (FVGP_OZAKI_TILE), best of reps; < 0 on error.  Allocates its own buffers. */
double fvgp_ozaki_i8_seconds(int64_t m, int64_t n, int64_t K, int tile, int reps, void* stream);
/* int8 multiply-accumulates launched by the INT8-slice path since the library was loaded (m n K per GEMM). */
unsigned long long fvgp_ozaki_mac_count(void);
int64_t fvgp_ozaki_work_bytes(int64_t m, int64_t n, int64_t k, int slices, int64_t nblock);
int fvgp_ozaki_gemm_nt(double* d_C, int64_t ldc, const double* d_A, int64_t lda, const double* d_B, int64_t ldb, int64_t m,
                       int64_t n, int64_t k, double sign, int lower, int64_t diag, int same_ab, int slices, int64_t nblock,
                       void* d_work, int64_t work_bytes, void* stream);

/* ---- measurement only: register-resident FP64 issue-rate probes (which: 0 = DMMA.8x8x4,
 * 1 = DFMA) giving the FP64 roofline denominator of the box.  d_scratch: 148*8*256 doubles. */
int fvgp_bench_fp64_peak(int which, int ctas_per_sm, int iters, double* d_scratch, double* h_tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FVGP_B200_H */
