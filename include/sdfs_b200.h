/*
 * sdfs_b200.h -- C ABI of the B200-native wealth-consumption-ratio / SDF solver.
 *
 * Drop-in boundary for ONE hot path of jstac/sdfs_via_autodiff: the discretised
 * operator  T w = 1 + beta * (H w^theta)^(1/theta),  H = diag(a_row) P diag(a_col),
 * the successive-approximation and Newton fixed-point loops around it, and the
 * SDF evaluated from the fixed point.  The reference has no FFI of its own (its
 * boundary is Python callables), so every entry point below cites the reference
 * function whose work it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 (SDFS_OK) or a negative SDFS_ERR_* code; the
 *     message is available through sdfs_last_error(); nothing throws or exits;
 *   - all tensors are C-contiguous fp64; pointers named d_* are DEVICE pointers
 *     on the context's GPU, h_* are host pointers;
 *   - one host thread per context; calls on one context are not re-entrant;
 *   - solver entry points run their whole loop on the device (no host sync per
 *     iteration) and return when the loop has finished.
 *   - there is no CPU implementation behind this ABI.
 */
#ifndef SDFS_B200_H
#define SDFS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFS_OK               0
#define SDFS_ERR_CUDA        -1
#define SDFS_ERR_ARG         -2
#define SDFS_ERR_NOMEM       -3
#define SDFS_ERR_UNSUPPORTED -4
#define SDFS_ERR_COMM        -5
#define SDFS_ERR_TIMEOUT     -6

#define SDFS_ABI_VERSION 1

typedef struct sdfs_ctx sdfs_ctx;       /* one GPU, one stream, one rank            */
typedef struct sdfs_op sdfs_op;         /* operator handle (dense or factor form)   */
typedef struct sdfs_factors sdfs_factors; /* discretised model: Markov factor arrays */

enum { SDFS_MODEL_SSY = 0, SDFS_MODEL_GCY = 1 };
enum { SDFS_KRYLOV_BICGSTAB = 0, SDFS_KRYLOV_GMRES = 1 };
enum { SDFS_STORAGE_DENSE = 0, SDFS_STORAGE_KRON = 1, SDFS_STORAGE_DENSE_REPLICATED = 2, SDFS_STORAGE_CONTINUOUS = 3,
       SDFS_STORAGE_KRON_LOCAL = 4 };

/* ---- context ---------------------------------------------------------- */
int sdfs_abi_version(void);
const char *sdfs_version_string(void);
/* device = CUDA ordinal. */
int sdfs_ctx_create(int device, sdfs_ctx **out);
int sdfs_ctx_destroy(sdfs_ctx *ctx);
/* ctx may be NULL: returns the message of the last failed call made without a
 * context (e.g. sdfs_ctx_create). */
const char *sdfs_last_error(sdfs_ctx *ctx);
int sdfs_ctx_sync(sdfs_ctx *ctx);
/* cudaDeviceSynchronize: used before consuming a foreign DLPack tensor. */
int sdfs_ctx_device_sync(sdfs_ctx *ctx);
int sdfs_ctx_device(sdfs_ctx *ctx, int *device, int *sm_count, size_t *free_bytes,
                    size_t *total_bytes);
/* Kernel launches issued by this context since creation (bench.py's
 * "gpu_launches" evidence). */
int64_t sdfs_ctx_launch_count(sdfs_ctx *ctx);
/* Per-launch CUDA-event timing of the dominant kernel (the dense row-stream pass) on the
 * launching stream: enable for up to max_launches launches (0 disables), then read the
 * summed device time and the number of launches recorded (the read synchronises). */
int sdfs_prof_enable(sdfs_ctx *ctx, int max_launches);
int sdfs_prof_read(sdfs_ctx *ctx, double *total_ms, int64_t *launches);
/* CUDA-event timer on the context's stream. */
int sdfs_timer_start(sdfs_ctx *ctx);
int sdfs_timer_stop_ms(sdfs_ctx *ctx, double *ms);

/* ---- memory (replaces jax.device_put, ssy_wc_ratio.py:227) -------------- */
int sdfs_malloc(sdfs_ctx *ctx, size_t bytes, void **d_ptr);
int sdfs_free(sdfs_ctx *ctx, void *d_ptr);
int sdfs_memset(sdfs_ctx *ctx, void *d_ptr, int value, size_t bytes);
int sdfs_h2d(sdfs_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int sdfs_d2h(sdfs_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int sdfs_d2d(sdfs_ctx *ctx, void *d_dst, const void *d_src, size_t bytes);
int sdfs_host_alloc_pinned(size_t bytes, void **h_ptr);
int sdfs_host_free_pinned(void *h_ptr);
int sdfs_fill_f64(sdfs_ctx *ctx, double *d_ptr, double value, int64_t n);

/* ---- discretisation on the device ---------------------------------------
 * Replaces discretize_ssy (ssy/discrete/ssy_wc_ratio.py:23-79) and
 * discretize_gcy (gcy/discrete/gcy_wc_ratio.py:31-131) including the
 * quantecon.rouwenhorst chains they call.  params/shapes use the reference's
 * orders: SSY params[13] = (beta,gamma,psi,mu_c,rho,phi_z,phi_c,rho_z,rho_c,
 * rho_lam,s_z,s_c,s_lam) (ssy_model.py:81), shapes[4] = (L,K,I,J);
 * GCY params[18] (gcy_model.py:72-75), shapes[6] =
 * (n_z,n_zpi,n_hz,n_hc,n_hzpi,n_hlam). */
int sdfs_factors_build(sdfs_ctx *ctx, int model, const double *h_params,
                       const int32_t *h_shapes, sdfs_factors **out);
/* Host-built factor arrays (the tuples the reference's discretisers return, in
 * their order: 10 arrays for SSY, 15 for GCY), copied to the device. */
int sdfs_factors_from_host(sdfs_ctx *ctx, int model, const double *h_params,
                           const int32_t *h_shapes, const double *const *h_arrays,
                           int n_arrays, sdfs_factors **out);
int sdfs_factors_destroy(sdfs_factors *f);
int sdfs_factors_count(sdfs_factors *f, int *n_arrays);
/* element count and device pointer of the idx-th array of the reference tuple */
int sdfs_factors_array(sdfs_factors *f, int idx, int64_t *n_elems, const double **d_ptr);

/* Log-linear closed form of the W/C ratio (wc_loglinear, ssy_model.py:143-153 /
 * gcy_model.py:146-157) evaluated at every grid state: h_coeffs[7] =
 * (A0, A_hlam, A_hc, A_hz, A_z, A_hzpi, A_zpi) (last two ignored for SSY);
 * d_out[n] = exp(value) if exponentiate else value.  Warm start for the solvers. */
int sdfs_factors_loglinear(sdfs_factors *f, const double *h_coeffs, int exponentiate, double *d_out);

/* ---- operators ----------------------------------------------------------
 * Dense single-index form (ssy/discrete/temp_ssy.py:49-106 P_x, :116-148 H,
 * :153-159 single_index_T).  d_P is row-major with leading dimension ld >= N
 * holding rows [row_begin,row_end) only (row-sharded ranks); the vectors have N
 * entries.  Pointers are borrowed: keep them alive until sdfs_op_destroy. */
int sdfs_op_from_dense(sdfs_ctx *ctx, const double *d_P, int64_t N, int64_t ld,
                       int64_t row_begin, int64_t row_end, const double *d_a_row,
                       const double *d_a_col, double beta, double theta, sdfs_op **out);
/* Expand the factors into a dense P (device kernel) or keep them in factor form
 * (storage = SDFS_STORAGE_KRON: sum-factorised apply, T_ssy ssy_wc_ratio.py:82-149,
 * T_gcy gcy_wc_ratio.py:134-236).  Dense storage honours the context's rank:
 * each rank materialises only its row slice; SDFS_STORAGE_DENSE_REPLICATED keeps the
 * full P on every rank (parameter sweeps shard columns, not rows).  Factor-form storage
 * honours the rank too where the leading axis can be contracted first (SSY, 9 <= shapes[0] <= 64,
 * shapes[0] >= ranks): each rank owns a slab of the leading axis, i.e. the contiguous rows
 * sdfs_op_info reports, reads the full input vector and exchanges result rows by peer stores;
 * other factor-form operators (GCY) and SDFS_STORAGE_KRON_LOCAL stay whole on every rank.
 * sdfs_op_info reports SDFS_STORAGE_KRON for both. */
int sdfs_op_from_factors(sdfs_ctx *ctx, sdfs_factors *f, int storage, sdfs_op **out);
/* Continuous-state operator ("next" row of the scope table): the T of
 * ssy/continuous_junnan/ssy_wc_ratio_continuous.py:125-226 and
 * gcy/continuous/gcy_wc_ratio_continuous.py -- conditional expectation by a shock rule (Gauss-Hermite
 * nodes/weights, or Monte-Carlo draws with weights 1/Q) and multilinear interpolation of w on uniform
 * grids (utils.py:6-23).  h_grids: the D grids concatenated (D = 4 SSY: h_lam,h_c,h_z,z; D = 6 GCY:
 * h_lam,h_c,h_z,h_zpi,z,z_pi); h_nodes: D x Q row-major; h_weights: Q.  The returned handle works
 * with sdfs_op_apply_T / _jvp and the three solvers. */
int sdfs_op_continuous(sdfs_ctx *ctx, int model, const double *h_params, const int32_t *h_sizes,
                       const double *h_grids, const double *h_nodes, const double *h_weights,
                       int64_t Q, sdfs_op **out);
/* lin_interp (utils.py:17-23; construct_wstar_callable, ssy_wc_ratio_continuous.py:304-326):
 * multilinear interpolation of d_vals (C-order tensor over D = 4 or 6 uniform grids, first value
 * h_g0[d], spacing h_intv[d]) at M points; d_x is D x M row-major; nearest-edge extension. */
int sdfs_interp_points(sdfs_ctx *ctx, int D, const int32_t *h_sizes, const double *h_g0,
                       const double *h_intv, const double *d_vals, const double *d_x, int64_t M,
                       double *d_out);
int sdfs_op_destroy(sdfs_op *op);
int sdfs_op_info(sdfs_op *op, int64_t *N, int64_t *ld, int64_t *row_begin,
                 int64_t *row_end, double *beta, double *theta, int *storage);
/* device pointers of the operator's own arrays (NULL when not materialised) */
int sdfs_op_arrays(sdfs_op *op, const double **d_P, const double **d_a_row,
                   const double **d_a_col, const double **d_e_sdf);
/* per-state SDF factor e_sdf(n) = exp(-gamma(mu_c+z(n)) + gamma^2 sigma_c(n)^2/2)
 * for operators that were not built from factors. */
int sdfs_op_set_esdf(sdfs_op *op, const double *d_e_sdf);
/* re-parametrise (gamma,psi,beta) of a factor-built operator without rebuilding P */
int sdfs_op_set_preferences(sdfs_op *op, double gamma, double psi, double beta);

/* T w  (T_ssy / T_gcy / single_index_T). d_w_in, d_w_out: N doubles. */
int sdfs_op_apply_T(sdfs_op *op, const double *d_w_in, double *d_w_out);
/* J_T(w) v = beta s^((1-theta)/theta) a_row P (a_col w^(theta-1) v)
 * (temp_ssy.py:204-216; replaces jax.jvp at solvers.py:87). */
int sdfs_op_apply_jvp(sdfs_op *op, const double *d_w, const double *d_v, double *d_out);
/* plain y = P x (used by tests and the CPU/GPU bandwidth comparison) */
int sdfs_op_apply_P(sdfs_op *op, const double *d_x, double *d_y);
/* diagnostic: `reps` back-to-back launches of the dense row-stream pass alone (mode 0 = T
 * epilogue, 3 = plain P x) on the x staged by the previous apply; average device ms. */
int sdfs_op_bench_pass(sdfs_op *op, int mode, int reps, double *avg_ms);
/* SDF from a fixed point (paper/autosdfs.tex:374-384; no reference code):
 * q_f(n) = sum_n' P(n,n') Mbar(n,n'),  euler(n) = beta^theta s(n)/(w(n)-1)^theta - 1.
 * Either output may be NULL. */
int sdfs_op_sdf(sdfs_op *op, const double *d_w, double *d_qf, double *d_euler);
/* explicit rows of Mbar(n, :) for n in h_rows[0..n_rows): d_out is n_rows x N */
int sdfs_op_sdf_rows(sdfs_op *op, const double *d_w, const int64_t *h_rows,
                     int64_t n_rows, double *d_out);

/* ---- solvers (device-resident loops) ------------------------------------
 * successive_approx (solvers.py:19-48): x <- T x until max|x_new-x| <= tol or
 * iters == max_iter; the iterate is accepted on the terminating step; a NaN
 * error ends the loop.  *iters = number of T evaluations.  d_err_hist (may be
 * NULL) receives the error of every iteration k with k % hist_stride == 0 at
 * index k / hist_stride (while < hist_cap) -- enough to reproduce the
 * reference's "iter = k, error = e" lines after the fact. */
int sdfs_solve_sa(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter,
                  double *d_w_out, int64_t *iters, double *final_err,
                  double *d_err_hist, int64_t hist_stride, int64_t hist_cap);
/* newton_solver (solvers.py:51-95): q(x) = x - J_g(x)^-1 g(x), g = T - id, fed to
 * the same successive-approximation rule.  Krylov = BICGSTAB reproduces
 * jax.scipy.sparse.linalg.bicgstab (x0 = 0, stop <r,r> <= max(rtol^2 <b,b>, atol^2),
 * early-exit half step, breakdown codes); GMRES is restarted GMRES(restart) with
 * the same stopping rule.  krylov_maxiter <= 0 means 10*N (JAX default).
 * h_outer_err[k] / h_inner_iters[k] (host arrays of length cap, may be NULL)
 * receive the step size and the Krylov iteration count of outer iteration k. */
int sdfs_solve_newton(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter,
                      int krylov, double rtol, double atol, int restart,
                      int64_t krylov_maxiter, double *d_w_out, int64_t *outer_iters,
                      double *final_err, double *h_outer_err, int64_t *h_inner_iters,
                      int64_t cap, int64_t *total_matvecs);

/* anderson_solver (solvers.py:98-124, a jaxopt.AndersonAcceleration wrapper with history_size=10,
 * mixing_frequency=4, beta=8.0, ridge=1e-6).  jaxopt is not vendored: its update rule is restated
 * (parity unpinned): history of the last m iterates/residuals, alpha from the ridge-regularised
 * bordered Gram system, extrapolation sum alpha_i (x_i + beta r_i) once the history is full and on
 * every mixing_frequency-th iteration, x <- T x otherwise; stops when ||T x - x||_2 <= tol. */
int sdfs_solve_anderson(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter,
                        int history_size, int mixing_frequency, double beta, double ridge,
                        double *d_w_out, int64_t *iters, double *final_err);

/* ---- batched (gamma, psi, beta) sweep -----------------------------------
 * B parameter columns share one P (P does not depend on preferences); each
 * step is S = P V (fp64 tensor-core GEMM) with per-column prologue/epilogue.
 * h_prefs is B x 3 (gamma, psi, beta).  d_W_out is N x B column-major
 * (column b = solution of parameter set b).  h_iters[B]. */
int sdfs_sweep_solve_sa(sdfs_op *op, const double *h_prefs, int64_t B, double w_init,
                        double tol, int64_t max_iter, double *d_W_out, int64_t *h_iters,
                        double *h_final_err);
/* Newton for every column at once: each column runs the reference's Newton iteration with its
 * own BiCGSTAB (JAX recurrence and stopping rule, rtol/atol as in sdfs_solve_newton); the
 * columns advance in lockstep so every Krylov mat-vec of all columns is one GEMM.
 * h_inner_total[b] = BiCGSTAB iterations summed over the outer iterations of column b. */
int sdfs_sweep_solve_newton(sdfs_op *op, const double *h_prefs, int64_t B, double w_init,
                            double tol, int64_t max_iter, double rtol, double atol,
                            int64_t krylov_maxiter, double *d_W_out, int64_t *h_outer_iters,
                            double *h_final_err, int64_t *h_inner_total, int64_t *total_gemms);
/* one batched T step on a resident panel (bench / tests): d_W_in, d_W_out N x B */
int sdfs_sweep_apply_T(sdfs_op *op, const double *h_prefs, int64_t B,
                       const double *d_W_in, double *d_W_out);
/* How the sweep forms S = P V for all columns:
 *   SDFS_SWEEP_DENSE  (0, default) the fp64 tensor-core GEMM against the stored P (BASELINE config 5);
 *   SDFS_SWEEP_FACTOR (1) the sum-factorised contraction over the Markov factors, batched over the
 *                     columns (16 N B bytes per mode instead of 2 N^2 B flop): same results to
 *                     rounding, same per-column iteration counts. */
#define SDFS_SWEEP_DENSE  0
#define SDFS_SWEEP_FACTOR 1
int sdfs_sweep_set_form(sdfs_op *op, int form);

/* ---- multi-GPU (one process per GPU) -------------------------------------
 * Row-sharded dense operators: rank g owns rows [N g/G, N (g+1)/G) of P and the
 * full vectors; one exchange of the result slices per application.  Every rank must issue
 * the same sequence of operator / solver calls (collective semantics).  Once the arenas are
 * mapped (sdfs_comm_arena_*), single-output applications (T, JVP, P x) and the solver loops
 * exchange through NVLink peer stores inside the kernel; without them, and for two-output
 * applications (SDF), the slices are all-gathered with NCCL.  A rank that never arrives makes
 * the others return SDFS_ERR_TIMEOUT after ~30 s instead of hanging.
 * sdfs_comm_unique_id fills a 128-byte NCCL id on rank 0; the host distributes
 * it (any transport) and every rank calls sdfs_comm_init. */
int sdfs_comm_unique_id(void *h_id128);
int sdfs_comm_init(sdfs_ctx *ctx, int rank, int nranks, const void *h_id128);
int sdfs_comm_rank(sdfs_ctx *ctx, int *rank, int *nranks);
int sdfs_comm_allgather_f64(sdfs_ctx *ctx, double *d_buf, int64_t count_per_rank);
int sdfs_comm_barrier(sdfs_ctx *ctx);
/* Peer-memory exchange (fused applications and solver loops): every rank exports a 64-byte
 * IPC handle of its exchange arena (sized for operators with N <= max_N), the host all-gathers
 * them, every rank maps its peers. */
int sdfs_comm_arena_export(sdfs_ctx *ctx, int64_t max_N, void *h_handle64);
int sdfs_comm_arena_import(sdfs_ctx *ctx, const void *h_handles64_all);

/* ---- DLPack --------------------------------------------------------------
 * Minimal producer/consumer for kDLCUDA float64 tensors (dlpack.h ABI v0.8
 * DLManagedTensor).  The Python side wraps these in PyCapsules. */
int sdfs_dlpack_export(sdfs_ctx *ctx, void *d_ptr, int ndim, const int64_t *shape,
                       void *owner_token, void (*release)(void *owner_token),
                       void **dl_managed_tensor);
int sdfs_dlpack_import(void *dl_managed_tensor, void **d_ptr, int *ndim,
                       int64_t *shape8, int *device, int64_t *n_elems);
void sdfs_dlpack_call_deleter(void *dl_managed_tensor);

#ifdef __cplusplus
}
#endif
#endif /* SDFS_B200_H */
