// Latency-bound regime: the reference's own default grids (SSY (2,3,4,5), N = 120).
// P (N^2 doubles <= 200 KB) lives in shared memory for the whole solve and one CTA
// iterates  w <- T w  with a single __syncthreads per iteration: no grid barrier, no
// global-memory round trip, no host involvement.  Same arithmetic and stopping rule as
// k_sa_loop (successive_approx, solvers.py:19-48).
#include "common.cuh"

#define SMALL_THREADS 1024
#define SMALL_WARPS 32
#define SMALL_MAX_ROWS_PER_WARP 8

struct SmallArgs {
    const double *P; int64_t ld; int N;
    const double *a_row, *a_col;
    double beta, theta;
    const double *w_init; double *w_out;
    double tol; long long max_iter;
    double *err_hist; long long hist_stride, hist_cap;
    long long *iters_out; double *final_err_out;
};

__global__ void __launch_bounds__(SMALL_THREADS, 1) k_sa_small(SmallArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int N = a.N;
    double *sP = sm;                          // N x N
    double *sx = sP + (size_t)N * N;          // 2 x N  (ping-pong matvec input)
    double *serr = sx + 2 * N;                // 2 x 32 (ping-pong per-warp sup-norms)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < N * N; e += SMALL_THREADS) sP[e] = a.P[(int64_t)(e / N) * a.ld + e % N];
    const double theta = a.theta, beta = a.beta, inv_theta = 1.0 / a.theta;
    // rows of this warp: warp, warp+32, ...; lane i (< nrows) owns the epilogue of row i
    int nrows = 0;
    for (int r = warp; r < N; r += SMALL_WARPS) ++nrows;
    const int my_row = warp + lane * SMALL_WARPS;
    const bool owner = lane < nrows;
    double w_mine = 0.0, ar = 0.0, ac = 0.0;
    if (owner) {
        w_mine = a.w_init[my_row];
        ar = a.a_row[my_row];
        ac = a.a_col[my_row];
        sx[my_row] = ac * pow(w_mine, theta);
    }
    if (threadIdx.x < 2 * SMALL_WARPS) serr[threadIdx.x] = 0.0;
    __syncthreads();
    long long it = 0;
    double error = a.tol + 1.0;
    while (error > a.tol && it < a.max_iter) {
        const int cur = (int)(it & 1), nxt = cur ^ 1;
        const double *x = sx + cur * N;
        double acc[SMALL_MAX_ROWS_PER_WARP];
#pragma unroll
        for (int i = 0; i < SMALL_MAX_ROWS_PER_WARP; ++i) acc[i] = 0.0;
        for (int c = lane; c < N; c += 32) {
            const double xc = x[c];
#pragma unroll
            for (int i = 0; i < SMALL_MAX_ROWS_PER_WARP; ++i)
                if (i < nrows) acc[i] = fma(sP[(size_t)(warp + i * SMALL_WARPS) * N + c], xc, acc[i]);
        }
        double s_mine = 0.0;
#pragma unroll
        for (int i = 0; i < SMALL_MAX_ROWS_PER_WARP; ++i)
            if (i < nrows) {
                const double s = warp_sum(acc[i]);
                if (lane == i) s_mine = s;
            }
        double d = 0.0;
        if (owner) {
            const double y = 1.0 + beta * pow(ar * s_mine, inv_theta);
            d = fabs(y - w_mine);
            w_mine = y;
            sx[nxt * N + my_row] = ac * pow(y, theta);
        }
        d = warp_nanmax(d);
        if (lane == 0) serr[cur * SMALL_WARPS + warp] = d;
        __syncthreads();
        error = warp_nanmax(serr[cur * SMALL_WARPS + lane]);
        if (threadIdx.x == 0 && a.err_hist && (it % a.hist_stride) == 0 && (it / a.hist_stride) < a.hist_cap)
            a.err_hist[it / a.hist_stride] = error;
        ++it;
    }
    if (owner) a.w_out[my_row] = w_mine;
    if (threadIdx.x == 0) {
        *a.iters_out = it;
        *a.final_err_out = error;
    }
}

// returns 1 if the small path handled the solve, 0 if not applicable, <0 on error
int small_sa_try(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter, double *d_w_out,
                 double *d_err_hist, int64_t hist_stride, int64_t hist_cap) {
    sdfs_ctx *ctx = op->ctx;
    if (op->storage != SDFS_STORAGE_DENSE) return 0;
    const DenseView &dv = op->dv;
    if (dv.row_begin != 0 || dv.row_end != dv.N) return 0;
    const int64_t N = dv.N;
    if (N > SMALL_WARPS * SMALL_MAX_ROWS_PER_WARP) return 0;
    const size_t smem = ((size_t)N * N + 2 * N + 2 * SMALL_WARPS) * sizeof(double);
    if (smem > 220 * 1024) return 0;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_sa_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SmallArgs a{};
    a.P = dv.P; a.ld = dv.ld; a.N = (int)N; a.a_row = dv.a_row; a.a_col = dv.a_col;
    a.beta = dv.beta; a.theta = dv.theta; a.w_init = d_w_init; a.w_out = d_w_out;
    a.tol = tol; a.max_iter = max_iter; a.err_hist = d_err_hist;
    a.hist_stride = hist_stride > 0 ? hist_stride : 1; a.hist_cap = d_err_hist ? hist_cap : 0;
    a.iters_out = (long long *)ctx->d_status;                 // LoopStatus.iters
    a.final_err_out = (double *)((char *)ctx->d_status + 32); // LoopStatus.final_err
    k_sa_small<<<1, SMALL_THREADS, smem, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 1;
}
