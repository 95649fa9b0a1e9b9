// Batched (gamma, psi, beta) parameter sweep (BASELINE config 5).
//
// B parameter sets share one transition matrix P (P depends on none of gamma, psi,
// beta), so one T step for all of them is a dense fp64 contraction
//     S[n, b] = sum_k P[n, k] V[b, k],   V[b, k] = exp(theta_b h_lam(k)) W[b, k]^theta_b
//     W'[b, n] = 1 + beta_b (a_row_b(n) S[n, b])^(1/theta_b)
// -- the only GEMM-shaped work on the path, and the only place tensor cores are used:
// mma.sync.m8n8k4 f64 (DMMA; tcgen05 has no f64 kind) fed from a 4-stage cp.async
// shared-memory ring.  Panels are stored column-major, i.e. [B][ldw] with every
// parameter column contiguous over the states, so both operands are K-contiguous and
// all fragment loads are conflict-free 8-byte LDS (row stride 20 doubles).
// The epilogue applies the per-column scalings, writes W', and folds the per-column
// sup-norm |W' - W| into one atomicMax per (warp, column); converged columns are frozen
// on the device, the host polls one flag every 64 steps.
#include "common.cuh"

#define TRY(x) do { int _rc = (x); if (_rc != SDFS_OK) return _rc; } while (0)

#define GM 128      // rows of P (states n) per CTA tile
#define GN 128      // parameter columns per CTA tile
#define GK 16       // K step
#define GS 20       // smem row stride in doubles (20 mod 16 == 4 -> conflict-free fragment loads)
#define GSTAGES 4
#define GTHREADS 256

struct SweepCols {            // per-column scalars, device arrays of length B
    const double *gamma, *theta, *beta;
    const int *done;          // 1 = converged earlier: column frozen
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool pred) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// V[b][k] = exp(theta_b * h_lam[k]) * W[b][k]^theta_b   (prologue; N*B elements)
__global__ void k_sweep_prologue(int64_t N, int64_t B, int64_t ldw, const double *__restrict__ h_lam,
                                 const double *__restrict__ W, SweepCols sc, double *__restrict__ V) {
    const int64_t total = N * B;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / N, k = e % N;
        const double th = sc.theta[b];
        V[b * ldw + k] = exp(th * h_lam[k]) * pow(W[b * ldw + k], th);
    }
}

// One batched T step.  grid = tiles_m * tiles_b (m-major so consecutive CTAs reuse the same
// rows of P out of L2).
__global__ void __launch_bounds__(GTHREADS, 1)
k_sweep_gemm(const double *__restrict__ P, int64_t N, int64_t ldp, const double *__restrict__ V, int64_t B,
             int64_t ldw, const double *__restrict__ sig_c, const double *__restrict__ mz,
             const double *__restrict__ W, double *__restrict__ Wn, SweepCols sc,
             unsigned long long *__restrict__ err_bits, int tiles_b) {
    extern __shared__ __align__(16) double smem[];
    double *sA = smem;                                   // [GSTAGES][GM][GS]
    double *sB = smem + (size_t)GSTAGES * GM * GS;       // [GSTAGES][GN][GS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;             // 4 x 2 warps, warp tile 32 x 64
    const int tile_m = blockIdx.x / tiles_b, tile_b = blockIdx.x % tiles_b;
    const int64_t m0 = (int64_t)tile_m * GM, b0 = (int64_t)tile_b * GN;
    const int64_t ksteps = (N + GK - 1) / GK;            // P and V are zero padded beyond N (ld multiple of 64)

    // loader mapping: 128 rows x 16 doubles = 128 x 8 chunks of 16 B per operand; 256 threads x 4 chunks
    auto load_stage = [&](int stage, int64_t ks) {
        const int64_t k0 = ks * GK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * GTHREADS;            // 0..1023
            const int row = c >> 3, seg = c & 7;         // seg: 2 doubles
            const int64_t n = m0 + row, b = b0 + row;
            const bool pa = n < N, pb = b < B;
            cp_async16(sA + ((size_t)stage * GM + row) * GS + seg * 2, P + (pa ? n : 0) * ldp + k0 + seg * 2, pa);
            cp_async16(sB + ((size_t)stage * GN + row) * GS + seg * 2, V + (pb ? b : 0) * ldw + k0 + seg * 2, pb);
        }
    };

    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
        if (s < ksteps) load_stage(s, s);
        cp_async_commit();
    }
    for (int64_t ks = 0; ks < ksteps; ++ks) {
        cp_async_wait<GSTAGES - 2>();
        __syncthreads();
        const int64_t nxt = ks + GSTAGES - 1;
        if (nxt < ksteps) load_stage((int)(nxt % GSTAGES), nxt);
        cp_async_commit();
        const int stage = (int)(ks % GSTAGES);
        const double *a_base = sA + ((size_t)stage * GM + wm * 32 + (lane >> 2)) * GS + (lane & 3);
        const double *b_base = sB + ((size_t)stage * GN + wn * 64 + (lane >> 2)) * GS + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_base[(size_t)i * 8 * GS + kk];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = b_base[(size_t)j * 8 * GS + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds S[n = m0 + wm*32 + i*8 + lane/4][b = b0 + wn*64 + j*8 + 2*(lane%4) + {0,1}]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t b = b0 + wn * 64 + j * 8 + 2 * (lane & 3) + h;
            double emax = 0.0;
            if (b < B) {
                const double g = sc.gamma[b], th = sc.theta[b], be = sc.beta[b];
                const int frozen = sc.done ? sc.done[b] : 0;
                const double omg = 1.0 - g, inv_th = 1.0 / th;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t n = m0 + wm * 32 + i * 8 + (lane >> 2);
                    if (n < N) {
                        const double t2 = omg * sig_c[n];
                        const double a_row = exp(0.5 * (t2 * t2)) * exp(omg * mz[n]);
                        const double w_old = W[b * ldw + n];
                        double y = 1.0 + be * pow(a_row * acc[i][j][h], inv_th);
                        if (frozen) y = w_old;
                        Wn[b * ldw + n] = y;
                        const double d = fabs(y - w_old);
                        emax = (d != d || emax != emax) ? d + emax : fmax(emax, d);   // NaN propagates
                    }
                }
            }
            // the 8 lanes with equal lane%4 hold the same column: combine, then one atomic per column
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                const double other = __shfl_xor_sync(0xffffffffu, emax, o);
                emax = (other != other || emax != emax) ? other + emax : fmax(emax, other);
            }
            if ((lane >> 2) == 0 && b < B && err_bits)
                atomicMax(err_bits + b, (unsigned long long)__double_as_longlong(fabs(emax)));
        }
    }
}

// after each step: count iterations, freeze converged columns, clear the error accumulators
__global__ void k_sweep_update(int64_t B, double tol, long long max_iter, unsigned long long *err_bits,
                               double *last_err, long long *iters, int *done, int *n_active) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (!done[b]) {
        const double e = __longlong_as_double((long long)err_bits[b]);
        iters[b] += 1;
        last_err[b] = e;
        if (!(e > tol) || iters[b] >= max_iter) {
            done[b] = 1;
            atomicSub(n_active, 1);
        }
    }
    err_bits[b] = 0ull;
}

__global__ void k_state_vectors(int model, KronView kv, const double *__restrict__ h_lam,
                                const double *__restrict__ sigma_c, const double *__restrict__ z, double mu_c,
                                double *__restrict__ hl, double *__restrict__ sc, double *__restrict__ mz) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < kv.N; n += (int64_t)gridDim.x * blockDim.x) {
        int c[SDFS_MAX_DIMS];
        int64_t rem = n;
        for (int d = kv.D - 1; d >= 0; --d) { c[d] = (int)(rem % kv.shape[d]); rem /= kv.shape[d]; }
        if (model == SDFS_MODEL_SSY) {
            hl[n] = h_lam[c[0]]; sc[n] = sigma_c[c[1]]; mz[n] = mu_c + z[c[2] * kv.shape[3] + c[3]];
        } else {
            hl[n] = h_lam[c[5]]; sc[n] = sigma_c[c[3]];
            mz[n] = mu_c + z[((c[1] * kv.shape[2] + c[2]) * kv.shape[4] + c[4]) * kv.shape[0] + c[0]];
        }
    }
}

struct SweepWork {
    double *hl = nullptr, *sc = nullptr, *mz = nullptr;     // per-state base vectors (N)
    double *gamma = nullptr, *theta = nullptr, *beta = nullptr, *last_err = nullptr;   // per column (B)
    unsigned long long *err_bits = nullptr;
    long long *iters = nullptr;
    int *done = nullptr, *n_active = nullptr;
    double *V = nullptr, *Wa = nullptr, *Wb = nullptr;      // panels [B][ldw]
    int64_t ldw = 0;
    void free_all() {
        void *ps[] = {hl, sc, mz, gamma, theta, beta, last_err, err_bits, iters, done, n_active, V, Wa, Wb};
        for (void *p : ps) if (p) cudaFree(p);
    }
};

static int sweep_setup(sdfs_op *op, const double *h_prefs, int64_t B, bool panels, SweepWork *w) {
    sdfs_ctx *ctx = op->ctx;
    if (op->storage != SDFS_STORAGE_DENSE || !op->factors)
        return sdfs_set_error(ctx, SDFS_ERR_ARG, "sweep needs a dense operator built from factors");
    if (op->dv.row_begin != 0 || op->dv.row_end != op->dv.N)
        return sdfs_set_error(ctx, SDFS_ERR_ARG, "sweep needs the full P on this rank (columns, not rows, are sharded)");
    const int64_t N = op->dv.N;
    w->ldw = round_up(N, 64);
    const sdfs_factors *f = op->factors;
    CUDA_TRY(ctx, cudaMalloc(&w->hl, N * 8)); CUDA_TRY(ctx, cudaMalloc(&w->sc, N * 8)); CUDA_TRY(ctx, cudaMalloc(&w->mz, N * 8));
    const double *h_lam = (f->model == SDFS_MODEL_SSY) ? f->d_arr[0] : f->d_arr[13];
    const double *sig_c = (f->model == SDFS_MODEL_SSY) ? f->d_arr[8] : f->d_arr[9];
    const double *z = (f->model == SDFS_MODEL_SSY) ? f->d_arr[6] : f->d_arr[0];
    k_state_vectors<<<(int)((N + 255) / 256 < 2048 ? (N + 255) / 256 : 2048), 256, 0, ctx->stream>>>(
        f->model, op->kv, h_lam, sig_c, z, op->mu_c, w->hl, w->sc, w->mz);
    ctx->launches++;
    std::vector<double> g(B), th(B), be(B);
    for (int64_t b = 0; b < B; ++b) {
        g[b] = h_prefs[3 * b]; const double psi = h_prefs[3 * b + 1]; be[b] = h_prefs[3 * b + 2];
        if (psi == 1.0 || g[b] == 1.0) return sdfs_set_error(ctx, SDFS_ERR_ARG, "column %lld: psi and gamma must differ from 1", (long long)b);
        th[b] = (1.0 - g[b]) / (1.0 - 1.0 / psi);
    }
    CUDA_TRY(ctx, cudaMalloc(&w->gamma, B * 8)); CUDA_TRY(ctx, cudaMalloc(&w->theta, B * 8)); CUDA_TRY(ctx, cudaMalloc(&w->beta, B * 8));
    CUDA_TRY(ctx, cudaMalloc(&w->last_err, B * 8)); CUDA_TRY(ctx, cudaMalloc(&w->err_bits, B * 8));
    CUDA_TRY(ctx, cudaMalloc(&w->iters, B * 8)); CUDA_TRY(ctx, cudaMalloc(&w->done, B * 4)); CUDA_TRY(ctx, cudaMalloc(&w->n_active, 4));
    CUDA_TRY(ctx, cudaMemcpyAsync(w->gamma, g.data(), B * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(w->theta, th.data(), B * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(w->beta, be.data(), B * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(w->last_err, 0, B * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(w->err_bits, 0, B * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(w->iters, 0, B * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(w->done, 0, B * 4, ctx->stream));
    const int nb = (int)B;
    CUDA_TRY(ctx, cudaMemcpyAsync(w->n_active, &nb, 4, cudaMemcpyHostToDevice, ctx->stream));
    const size_t pbytes = (size_t)B * w->ldw * 8;
    CUDA_TRY(ctx, cudaMalloc(&w->V, pbytes));
    CUDA_TRY(ctx, cudaMemsetAsync(w->V, 0, pbytes, ctx->stream));
    if (panels) {
        CUDA_TRY(ctx, cudaMalloc(&w->Wa, pbytes)); CUDA_TRY(ctx, cudaMalloc(&w->Wb, pbytes));
        CUDA_TRY(ctx, cudaMemsetAsync(w->Wa, 0, pbytes, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(w->Wb, 0, pbytes, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));    // host vectors g/th/be go out of scope
    return SDFS_OK;
}

static int sweep_step(sdfs_op *op, SweepWork &w, int64_t B, const double *Win, double *Wout, bool track) {
    sdfs_ctx *ctx = op->ctx;
    const int64_t N = op->dv.N;
    SweepCols sc{w.gamma, w.theta, w.beta, track ? w.done : nullptr};
    const int64_t tot = N * B;
    k_sweep_prologue<<<(int)((tot + 255) / 256 < (int64_t)ctx->sm_count * 16 ? (tot + 255) / 256 : (int64_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        N, B, w.ldw, w.hl, Win, sc, w.V);
    const int tiles_m = (int)((N + GM - 1) / GM), tiles_b = (int)((B + GN - 1) / GN);
    const size_t smem = (size_t)GSTAGES * (GM + GN) * GS * sizeof(double);
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
    if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    k_sweep_gemm<<<tiles_m * tiles_b, GTHREADS, smem, ctx->stream>>>(op->dv.P, N, op->dv.ld, w.V, B, w.ldw, w.sc, w.mz, Win,
                                                                     Wout, sc, track ? w.err_bits : nullptr, tiles_b);
    if (prof) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

__global__ void k_fill_panel(double *W, int64_t N, int64_t B, int64_t ldw, double v) {
    const int64_t tot = N * B;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x)
        W[(e / N) * ldw + e % N] = v;
}
__global__ void k_copy_panel(const double *src, int64_t lds, double *dst, int64_t ldd, int64_t N, int64_t B) {
    const int64_t tot = N * B;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x)
        dst[(e / N) * ldd + e % N] = src[(e / N) * lds + e % N];
}

extern "C" {

int sdfs_sweep_apply_T(sdfs_op *op, const double *h_prefs, int64_t B, const double *d_W_in, double *d_W_out) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_sweep_apply_T: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, h_prefs && d_W_in && d_W_out && B >= 1);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    SweepWork w;
    int rc = sweep_setup(op, h_prefs, B, true, &w);
    const int64_t N = op->dv.N;
    const int grid = ctx->sm_count * 8;
    if (rc == SDFS_OK) {
        k_copy_panel<<<grid, 256, 0, ctx->stream>>>(d_W_in, N, w.Wa, w.ldw, N, B);
        rc = sweep_step(op, w, B, w.Wa, w.Wb, false);
    }
    if (rc == SDFS_OK) {
        k_copy_panel<<<grid, 256, 0, ctx->stream>>>(w.Wb, w.ldw, d_W_out, N, N, B);
        ctx->launches += 2;
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = sdfs_set_error(ctx, SDFS_ERR_CUDA, "sweep apply: %s", cudaGetErrorString(e));
    }
    w.free_all();
    return rc;
}

int sdfs_sweep_solve_sa(sdfs_op *op, const double *h_prefs, int64_t B, double w_init, double tol, int64_t max_iter,
                        double *d_W_out, int64_t *h_iters, double *h_final_err) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_sweep_solve_sa: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, h_prefs && d_W_out && B >= 1 && max_iter >= 0);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    SweepWork w;
    int rc = sweep_setup(op, h_prefs, B, true, &w);
    const int64_t N = op->dv.N;
    const int grid = ctx->sm_count * 8;
    double *cur = w.Wa, *nxt = w.Wb;
    if (rc == SDFS_OK) {
        k_fill_panel<<<grid, 256, 0, ctx->stream>>>(cur, N, B, w.ldw, w_init);
        ctx->launches++;
        int active = (int)B;
        int64_t steps = 0;
        while (rc == SDFS_OK && active > 0 && steps < max_iter) {
            // a burst of steps with no host synchronisation; convergence is tracked on the device
            for (int i = 0; i < 64 && steps < max_iter && rc == SDFS_OK; ++i, ++steps) {
                rc = sweep_step(op, w, B, cur, nxt, true);
                k_sweep_update<<<(int)((B + 127) / 128), 128, 0, ctx->stream>>>(B, tol, (long long)max_iter, w.err_bits, w.last_err,
                                                                                w.iters, w.done, w.n_active);
                ctx->launches++;
                double *t = cur; cur = nxt; nxt = t;
            }
            if (rc != SDFS_OK) break;
            cudaError_t e = cudaMemcpyAsync(&active, w.n_active, 4, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) rc = sdfs_set_error(ctx, SDFS_ERR_CUDA, "sweep solve: %s", cudaGetErrorString(e));
        }
    }
    if (rc == SDFS_OK) {
        k_copy_panel<<<grid, 256, 0, ctx->stream>>>(cur, w.ldw, d_W_out, N, N, B);
        ctx->launches++;
        std::vector<long long> it(B);
        cudaError_t e = cudaMemcpyAsync(it.data(), w.iters, B * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && h_final_err) e = cudaMemcpyAsync(h_final_err, w.last_err, B * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = sdfs_set_error(ctx, SDFS_ERR_CUDA, "sweep solve: %s", cudaGetErrorString(e));
        if (h_iters) for (int64_t b = 0; b < B; ++b) h_iters[b] = it[b];
    }
    w.free_all();
    return rc;
}

}  // extern "C"
