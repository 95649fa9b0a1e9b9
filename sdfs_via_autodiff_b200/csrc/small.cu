// Latency-bound regime: the reference's own default grids (SSY (2,3,4,5), N = 120).
// P (N^2 doubles <= 200 KB) lives in shared memory for the whole solve and one CTA
// iterates  w <- T w  with a single __syncthreads per iteration: no grid barrier, no
// global-memory round trip, no host involvement.  Same arithmetic and stopping rule as
// k_sa_loop (successive_approx, solvers.py:19-48).
#include "common.cuh"
#include "rowdot.cuh"

#include <stdlib.h>
#define SMALL_MAX_N 160
#define TRY_RC(x) do { int _rc = (x); if (_rc != SDFS_OK) return _rc; } while (0)

struct SmallArgs {
    const double *P; int64_t ld; int N;
    const double *a_row, *a_col;
    double beta, theta;
    const double *w_init; double *w_out;
    double tol; long long max_iter;
    double *err_hist; long long hist_stride, hist_cap;
    long long *iters_out; double *final_err_out;
};

// x^e for x > 0 through exp(e log x): one log and one exp instead of pow's extended-precision
// path.  |e log x| <= ~250 here, so the relative error is <= ~250 ulp(1) ~ 3e-14 in w^theta and
// shrinks by the factor 1/|theta| when T takes the 1/theta power: far inside the 1e-10 contract.
// NaN for x < 0 and inf for x = 0, e < 0, like pow.
template <int MODE>
__device__ __forceinline__ double pw(double x, double e) {
    return MODE == 0 ? pow(x, e) : exp(e * log(x));
}

template <int SMALL_WARPS, int MODE>
__global__ void __launch_bounds__(SMALL_WARPS * 32, 1) k_sa_small(SmallArgs a) {
    constexpr int SMALL_THREADS = SMALL_WARPS * 32;
    constexpr int SMALL_MAX_ROWS_PER_WARP = (SMALL_MAX_N + SMALL_WARPS - 1) / SMALL_WARPS;
    extern __shared__ __align__(16) double sm[];
    const int N = a.N;
    double *sP = sm;                          // N x N
    double *sx = sP + (size_t)N * N;          // 2 x N  (ping-pong matvec input)
    double *serr = sx + 2 * N;                // 2 x 32 (ping-pong per-warp sup-norms; unused tail = 0)
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
    if (threadIdx.x < 64) serr[threadIdx.x] = 0.0;
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
            const double y = 1.0 + beta * pw<MODE>(ar * s_mine, inv_theta);
            d = fabs(y - w_mine);
            w_mine = y;
            sx[nxt * N + my_row] = ac * pw<MODE>(y, theta);
        }
        d = warp_nanmax(d);
        if (lane == 0) serr[cur * 32 + warp] = d;
        __syncthreads();
        error = warp_nanmax(serr[cur * 32 + lane]);
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

// ---------------------------------------------------------------------------
// N <= 128: register-resident P.  512 threads; 4 lanes per row, 8 rows per warp; lane l4 of a
// row holds the 32 matrix entries of columns {8k + 2 l4, 8k + 2 l4 + 1}, k = 0..15, in
// registers for the whole solve, so an iteration's mat-vec is 16 broadcast LDS.128 of x, 32
// DFMA and 2 shuffle stages per lane.  Measured B200 latencies (tools/lat_probe.cu: pow 844
// cycles for one warp but 2116 when 32 warps issue it, warp_sum(f64) ~200, __syncthreads
// ~75) drive the rest: the transcendental epilogue runs only in the first N threads (4 full
// warps, exp/log form), the sup-norm uses two REDUX.MAX on the high/low words of |dy|
// instead of five 64-bit shuffle stages, and the four warp maxima meet in shared memory.
// Two __syncthreads per iteration.
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_sa_small_reg(SmallArgs a) {
    __shared__ __align__(16) double sx[2][128];
    __shared__ double ss[128];
    __shared__ unsigned long long smax[2][4];
    const int N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l4 = lane & 3;
    const int row = warp * 8 + (lane >> 2);
    const bool row_ok = row < N;
    double2 pr[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int c = 8 * k + 2 * l4;
        pr[k].x = (row_ok && c < N) ? a.P[(int64_t)row * a.ld + c] : 0.0;
        pr[k].y = (row_ok && c + 1 < N) ? a.P[(int64_t)row * a.ld + c + 1] : 0.0;
    }
    const double theta = a.theta, beta = a.beta, inv_theta = 1.0 / a.theta;
    const bool owner = tid < N;
    double w_mine = 0.0, ar = 0.0, ac = 0.0;
    if (tid < 128) { sx[0][tid] = 0.0; sx[1][tid] = 0.0; }
    if (tid < 8) smax[tid >> 2][tid & 3] = 0ull;
    __syncthreads();
    if (owner) {
        w_mine = a.w_init[tid];
        ar = a.a_row[tid];
        ac = a.a_col[tid];
        sx[0][tid] = ac * pow(w_mine, theta);
    }
    __syncthreads();
    long long it = 0;
    double error = a.tol + 1.0;
    while (error > a.tol && it < a.max_iter) {
        const int cur = (int)(it & 1), nxt = cur ^ 1;
        const double *x = sx[cur];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            const double2 x0 = *reinterpret_cast<const double2 *>(x + 8 * k + 2 * l4);
            const double2 x1 = *reinterpret_cast<const double2 *>(x + 8 * (k + 1) + 2 * l4);
            a0 = fma(pr[k].x, x0.x, a0);
            a1 = fma(pr[k].y, x0.y, a1);
            a2 = fma(pr[k + 1].x, x1.x, a2);
            a3 = fma(pr[k + 1].y, x1.y, a3);
        }
        double s = (a0 + a1) + (a2 + a3);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (l4 == 0 && row_ok) ss[row] = s;
        __syncthreads();
        if (warp < 4) {
            double d = 0.0;
            if (owner) {
                const double y = 1.0 + beta * pw<MODE>(ar * ss[tid], inv_theta);
                d = fabs(y - w_mine);
                w_mine = y;
                sx[nxt][tid] = ac * pw<MODE>(y, theta);
            }
            // NaN-propagating max of non-negative doubles through their bit patterns
            const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
            const int hi = (int)(bits >> 32);
            const int mhi = __reduce_max_sync(0xffffffffu, hi);
            const unsigned lo = (hi == mhi) ? (unsigned)bits : 0u;
            const unsigned mlo = __reduce_max_sync(0xffffffffu, lo);
            if (lane == 0) smax[cur][warp] = ((unsigned long long)(unsigned)mhi << 32) | mlo;
        }
        __syncthreads();
        unsigned long long m = smax[cur][0];
        m = max(m, smax[cur][1]);
        m = max(m, smax[cur][2]);
        m = max(m, smax[cur][3]);
        error = __longlong_as_double((long long)m);
        if (tid == 0 && a.err_hist && (it % a.hist_stride) == 0 && (it / a.hist_stride) < a.hist_cap)
            a.err_hist[it / a.hist_stride] = error;
        ++it;
    }
    if (owner) a.w_out[tid] = w_mine;
    if (tid == 0) {
        *a.iters_out = it;
        *a.final_err_out = error;
    }
}

// ---------------------------------------------------------------------------
// Mid-size grids in factor form (the reference's default GCY grid, N = 3^6 = 729, lives here): the whole
// state vector AND every factor matrix stay in the shared memory of ONE CTA for the whole solve.  An
// iteration is D mode contractions between two shared-memory vectors (thread per output element, the
// element -> (matrix row, fibre base) maps decoded once before the loop), the epilogue fused with the next
// iteration's prologue, and a block-wide sup-norm: D + 2 __syncthreads per iteration instead of one grid
// barrier per mode (k_sa_loop<KronLoopOp> costs 25-50 us per iteration on these grids, all of it barrier and
// staging latency).  Same stopping rule as successive_approx (solvers.py:19-48): the error is taken before
// the update is accepted.
// ---------------------------------------------------------------------------
#define KSMALL_THREADS 512
struct KronSmallArgs {
    const double *w_init; double *w_out;
    double *wbuf;                  // 2 x ldn doubles of global scratch: ping-pong iterates (stay in L1/L2)
    int *tab;                      // [n_modes][2][ldn] ints of global scratch: fibre base, matrix-row offset per element
    long long ldn;
    int poff[SDFS_MAX_DIMS];       // offset of each mode's matrices in the shared-memory pool
    double tol; long long max_iter;
    double *err_hist; long long hist_stride, hist_cap;
    long long *iters_out; double *final_err_out;
};

__global__ void __launch_bounds__(KSMALL_THREADS, 1)
k_sa_kron_small(const __grid_constant__ KronView kv, const __grid_constant__ KronSmallArgs a) {
    extern __shared__ __align__(16) double ksm[];
    __shared__ double s_red[KSMALL_THREADS / 32];
    const int N = (int)kv.N;
    const int ldn = (int)a.ldn;
    double *va = ksm, *vb = ksm + ldn, *pool = ksm + 2 * ldn;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double theta = kv.theta, beta = kv.beta, inv_theta = 1.0 / kv.theta;
    double *w_prev = a.wbuf, *w_next = a.wbuf + a.ldn;
    // factor matrices -> shared memory (contiguous [n_mats][n][n] per mode)
    for (int m = 0; m < kv.n_modes; ++m) {
        const KronMode &md = kv.modes[m];
        const int n = kv.shape[md.dim];
        const int cnt = (int)md.Mcount * n * n;
        for (int e = tid; e < cnt; e += KSMALL_THREADS) pool[a.poff[m] + e] = md.mat[e];
    }
    // element maps, once: for output element idx of mode m, the base of its fibre and its matrix row
    for (int idx = tid; idx < N; idx += KSMALL_THREADS) {
        int c[SDFS_MAX_DIMS], rem = idx;
        for (int d = kv.D - 1; d >= 0; --d) { c[d] = rem % kv.shape[d]; rem /= kv.shape[d]; }
        for (int m = 0; m < kv.n_modes; ++m) {
            const KronMode &md = kv.modes[m];
            const int n = kv.shape[md.dim];
            int mat = 0;
            for (int d = 0; d < kv.D; ++d) mat += c[d] * md.mstride[d];
            a.tab[(2 * m) * ldn + idx] = idx - c[md.dim] * (int)md.stride;
            a.tab[(2 * m + 1) * ldn + idx] = a.poff[m] + (mat * n + c[md.dim]) * n;
        }
        const double w0 = a.w_init[idx];
        w_prev[idx] = w0;
        va[idx] = kv.a_col[idx] * pow_pos(w0, theta);
    }
    __syncthreads();
    long long it = 0;
    double error = a.tol + 1.0;
    double *in = va, *out = vb;
    while (error > a.tol && it < a.max_iter) {
        for (int m = 0; m < kv.n_modes; ++m) {
            const int n = kv.shape[kv.modes[m].dim];
            const int stride = (int)kv.modes[m].stride;
            const int *tb = a.tab + (2 * m) * ldn, *tr = tb + ldn;
            for (int i0 = tid; i0 < N; i0 += 2 * KSMALL_THREADS) {           // two outputs per trip: independent FMA chains
                const int i1 = i0 + KSMALL_THREADS;
                const bool v1 = i1 < N;
                const double *ra = pool + tr[i0], *xa = in + tb[i0];
                const double *rb = pool + (v1 ? tr[i1] : tr[i0]), *xb = in + (v1 ? tb[i1] : tb[i0]);
                double sa0 = 0.0, sa1 = 0.0, sb0 = 0.0, sb1 = 0.0;
                int j = 0;
                for (; j + 1 < n; j += 2) {
                    sa0 = fma(ra[j], xa[j * stride], sa0);
                    sa1 = fma(ra[j + 1], xa[(j + 1) * stride], sa1);
                    sb0 = fma(rb[j], xb[j * stride], sb0);
                    sb1 = fma(rb[j + 1], xb[(j + 1) * stride], sb1);
                }
                if (j < n) { sa0 = fma(ra[j], xa[j * stride], sa0); sb0 = fma(rb[j], xb[j * stride], sb0); }
                out[i0] = sa0 + sa1;
                if (v1) out[i1] = sb0 + sb1;
            }
            __syncthreads();
            double *t = in; in = out; out = t;
        }
        // epilogue + next prologue on the finished contraction (in), element by element in place
        double e = 0.0;
        for (int n0 = tid; n0 < N; n0 += 2 * KSMALL_THREADS) {              // two independent log/exp chains per trip
            const int n1 = n0 + KSMALL_THREADS;
            const bool v1 = n1 < N;
            const double sa = kv.a_row[n0] * in[n0], sb = v1 ? kv.a_row[n1] * in[n1] : 1.0;
            const double ya = 1.0 + beta * pow_pos(sa, inv_theta), yb = 1.0 + beta * pow_pos(sb, inv_theta);
            const double xa = kv.a_col[n0] * pow_pos(ya, theta), xb = v1 ? kv.a_col[n1] * pow_pos(yb, theta) : 0.0;
            e = nanmax(e, fabs(ya - w_prev[n0]));
            w_next[n0] = ya;
            in[n0] = xa;
            if (v1) {
                e = nanmax(e, fabs(yb - w_prev[n1]));
                w_next[n1] = yb;
                in[n1] = xb;
            }
        }
        e = warp_nanmax(e);
        if (lane == 0) s_red[warp] = e;
        __syncthreads();
        double mx = s_red[0];
#pragma unroll
        for (int w = 1; w < KSMALL_THREADS / 32; ++w) mx = nanmax(mx, s_red[w]);
        error = mx;
        __syncthreads();                           // s_red is rewritten next iteration; the staged vector is complete
        if (tid == 0 && a.err_hist && (it % a.hist_stride) == 0 && (it / a.hist_stride) < a.hist_cap)
            a.err_hist[it / a.hist_stride] = error;
        double *t = w_prev; w_prev = w_next; w_next = t;
        ++it;
    }
    for (int n = tid; n < N; n += KSMALL_THREADS) a.w_out[n] = w_prev[n];
    if (tid == 0) {
        *a.iters_out = it;
        *a.final_err_out = error;
    }
}

static int small_sa_kron_try(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter, double *d_w_out,
                             double *d_err_hist, int64_t hist_stride, int64_t hist_cap) {
    sdfs_ctx *ctx = op->ctx;
    const KronView &kv = op->kv;
    static const long long max_n = getenv("SDFS_SMALL_KRON_MAX") ? atoll(getenv("SDFS_SMALL_KRON_MAX")) : 8192;
    if (kv.N > max_n || op->kron_sharded) return 0;
    KronSmallArgs a{};
    long long pool = 0;
    for (int m = 0; m < kv.n_modes; ++m) {
        const int n = kv.shape[kv.modes[m].dim];
        a.poff[m] = (int)pool;
        pool += kv.modes[m].Mcount * n * n;
    }
    const long long ldn = (kv.N + 1) & ~1LL;
    const size_t smem = (size_t)(2 * ldn + pool) * sizeof(double);
    if (smem > 200 * 1024) return 0;
    // global scratch from the work vectors: 2 iterates + 2 int tables per mode (one double holds two ints)
    TRY_RC(op_ensure_work(op, 16));
    if (ldn > op->ldv || 2 + kv.n_modes > 16) return 0;
    a.w_init = d_w_init; a.w_out = d_w_out;
    a.wbuf = op->work; a.ldn = ldn;
    a.tab = (int *)(op->work + 2 * op->ldv);
    a.tol = tol; a.max_iter = max_iter; a.err_hist = d_err_hist;
    a.hist_stride = hist_stride > 0 ? hist_stride : 1; a.hist_cap = d_err_hist ? hist_cap : 0;
    a.iters_out = (long long *)ctx->d_status;
    a.final_err_out = (double *)((char *)ctx->d_status + 32);
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_sa_kron_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sa_kron_small<<<1, KSMALL_THREADS, smem, ctx->stream>>>(kv, a);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 1;
}

// returns 1 if the small path handled the solve, 0 if not applicable, <0 on error
int small_sa_try(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter, double *d_w_out,
                 double *d_err_hist, int64_t hist_stride, int64_t hist_cap) {
    sdfs_ctx *ctx = op->ctx;
    if (op->storage == SDFS_STORAGE_KRON)
        return small_sa_kron_try(op, d_w_init, tol, max_iter, d_w_out, d_err_hist, hist_stride, hist_cap);
    if (op->storage != SDFS_STORAGE_DENSE) return 0;
    const DenseView &dv = op->dv;
    if (dv.row_begin != 0 || dv.row_end != dv.N) return 0;
    const int64_t N = dv.N;
    // a dense operator that was built from factors still has them: between 161 and 8192 states successive
    // approximation is fastest as the one-CTA factor-form kernel (5.5 us per iteration at 729 states against
    // 7 us per iteration for the dense cooperative loop), while Newton on the same operator uses the dense pass
    // (18 us per application against 54 us for six factor-form modes) - storage "auto" relies on this
    if (N > SMALL_MAX_N && op->factors != nullptr && op->kv.n_modes >= 2)
        return small_sa_kron_try(op, d_w_init, tol, max_iter, d_w_out, d_err_hist, hist_stride, hist_cap);
    if (N > SMALL_MAX_N) return 0;
    const size_t smem = ((size_t)N * N + 2 * N + 64) * sizeof(double);
    if (smem > 220 * 1024) return 0;
    const char *ev = getenv("SDFS_SMALL_VARIANT");     // experiment switch: "<warps><mode>", e.g. 320, 321, 81
    const int variant = ev ? atoi(ev) : 321;
    SmallArgs a{};
    a.P = dv.P; a.ld = dv.ld; a.N = (int)N; a.a_row = dv.a_row; a.a_col = dv.a_col;
    a.beta = dv.beta; a.theta = dv.theta; a.w_init = d_w_init; a.w_out = d_w_out;
    a.tol = tol; a.max_iter = max_iter; a.err_hist = d_err_hist;
    a.hist_stride = hist_stride > 0 ? hist_stride : 1; a.hist_cap = d_err_hist ? hist_cap : 0;
    a.iters_out = (long long *)ctx->d_status;                 // LoopStatus.iters
    a.final_err_out = (double *)((char *)ctx->d_status + 32); // LoopStatus.final_err
#define LAUNCH_SMALL(W, M)                                                                                   \
    do {                                                                                                     \
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_sa_small<W, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_sa_small<W, M><<<1, W * 32, smem, ctx->stream>>>(a);                                               \
    } while (0)
    if (N <= 128 && (variant == 321 || variant == 1)) {          // default: register-resident P
        k_sa_small_reg<1><<<1, 512, 0, ctx->stream>>>(a);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        return 1;
    }
    if (N <= 128 && variant == 0) {                              // register-resident P, pow()
        k_sa_small_reg<0><<<1, 512, 0, ctx->stream>>>(a);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        return 1;
    }
    switch (variant) {
        case 320: LAUNCH_SMALL(32, 0); break;
        case 80: LAUNCH_SMALL(8, 0); break;
        case 81: LAUNCH_SMALL(8, 1); break;
        case 161: LAUNCH_SMALL(16, 1); break;
        default: LAUNCH_SMALL(32, 1); break;
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return 1;
}
