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
#include "rowdot.cuh"      // factor-form mode contractions (fused sweep kernel)

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

// GEMM epilogue modes
//  0: W' = 1 + beta (a_row S)^(1/theta), per-column sup-norm |W' - W|          (SA step)
//  1: G = (1 + beta (a_row S)^(1/theta)) - W ; D = beta a_row (a_row S)^((1-theta)/theta)   (Newton residual)
//  2: out = D .* S - Vsub                                                     (J_g v, Krylov mat-vec)
struct SweepEpi {
    int mode;
    const double *W;          // modes 0, 1
    double *out0;             // W' | G | J_g v
    double *out1;             // mode 1: D
    const double *D, *Vsub;   // mode 2
    unsigned long long *err_bits;   // mode 0
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src, bool pred) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// V[b][k] = exp(theta_b * h_lam[k]) * W[b][k]^theta_b   (prologue; N*B elements)
__global__ void k_sweep_prologue(int64_t N, int64_t B, int64_t ldw, const double *__restrict__ h_lam,
                                 const double *__restrict__ W, SweepCols sc, double *__restrict__ V) {
    const int64_t total = N * B;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / N, k = e % N;
        const double th = sc.theta[b];
        // exp(theta h) w^theta as one exponential: 1 log + 1 exp instead of exp + pow (relative error
        // <= |theta (h + log w)| ulp ~ 1e-14, shrunk again by the 1/theta power of the epilogue)
        V[b * ldw + k] = exp(th * (h_lam[k] + log(W[b * ldw + k])));
    }
}

// One batched T step.  grid = tiles_m * tiles_b (m-major so consecutive CTAs reuse the same
// rows of P out of L2).
template <int TN>      // parameter columns per CTA tile: 128 (warp tile 32 x 64) or 64 (warp tile 32 x 32)
__global__ void __launch_bounds__(GTHREADS, 1)
k_sweep_gemm(const double *__restrict__ P, int64_t N, int64_t ldp, const double *__restrict__ V, int64_t B,
             int64_t ldw, const double *__restrict__ sig_c, const double *__restrict__ mz,
             SweepEpi ep, SweepCols sc, int tiles_b) {
    extern __shared__ __align__(16) double smem[];
    double *sA = smem;                                   // [GSTAGES][GM][GS]
    double *sB = smem + (size_t)GSTAGES * GM * GS;       // [GSTAGES][TN][GS]
    constexpr int NJ = TN / 16;                          // 8x8 column blocks per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1;             // 4 x 2 warps, warp tile 32 x (TN/2)
    const int tile_m = blockIdx.x / tiles_b, tile_b = blockIdx.x % tiles_b;
    const int64_t m0 = (int64_t)tile_m * GM, b0 = (int64_t)tile_b * TN;
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
            if (row < TN)
                cp_async16(sB + ((size_t)stage * TN + row) * GS + seg * 2, V + (pb ? b : 0) * ldw + k0 + seg * 2, pb);
        }
    };

    double acc[4][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

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
        const double *b_base = sB + ((size_t)stage * TN + wn * (TN / 2) + (lane >> 2)) * GS + (lane & 3);
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            double af[4], bf[NJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a_base[(size_t)i * 8 * GS + kk];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bf[j] = b_base[(size_t)j * 8 * GS + kk];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds S[n = m0 + wm*32 + i*8 + lane/4][b = b0 + wn*(TN/2) + j*8 + 2*(lane%4) + {0,1}]
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t b = b0 + wn * (TN / 2) + j * 8 + 2 * (lane & 3) + h;
            double emax = 0.0;
            if (b < B) {
                if (ep.mode == 2) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t n = m0 + wm * 32 + i * 8 + (lane >> 2);
                        if (n < N) ep.out0[b * ldw + n] = ep.D[b * ldw + n] * acc[i][j][h] - ep.Vsub[b * ldw + n];
                    }
                } else {
                    const double g = sc.gamma[b], th = sc.theta[b], be = sc.beta[b];
                    const int frozen = sc.done ? sc.done[b] : 0;
                    const double omg = 1.0 - g, inv_th = 1.0 / th;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int64_t n = m0 + wm * 32 + i * 8 + (lane >> 2);
                        if (n < N) {
                            const double t2 = omg * sig_c[n];
                            const double a_row = exp(0.5 * (t2 * t2)) * exp(omg * mz[n]);
                            const double w_old = ep.W[b * ldw + n];
                            const double sv = a_row * acc[i][j][h];
                            double y = 1.0 + be * pow(sv, inv_th);
                            if (ep.mode == 0) {
                                if (frozen) y = w_old;
                                ep.out0[b * ldw + n] = y;
                                const double d = fabs(y - w_old);
                                emax = (d != d || emax != emax) ? d + emax : fmax(emax, d);   // NaN propagates
                            } else {
                                ep.out0[b * ldw + n] = y - w_old;
                                ep.out1[b * ldw + n] = be * pow(sv, (1.0 - th) * inv_th) * a_row;
                            }
                        }
                    }
                }
            }
            if (ep.mode == 0) {
                // the 8 lanes with equal lane%4 hold the same column: combine, then one atomic per column
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    const double other = __shfl_xor_sync(0xffffffffu, emax, o);
                    emax = (other != other || emax != emax) ? other + emax : fmax(emax, other);
                }
                if ((lane >> 2) == 0 && b < B && ep.err_bits)
                    atomicMax(ep.err_bits + b, (unsigned long long)__double_as_longlong(fabs(emax)));
            }
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

// helpers of the register-blocked short-axis contraction (kron_mode_fibre2 below)
static __host__ __device__ inline int kron_mode_nmats(const KronMode &md) {
    int nm = 1;
    for (int a = 0; a < md.nM; ++a) nm += (md.Mshape[a] - 1) * md.Mmat[a];
    return nm;
}
static __host__ __device__ inline int fibre2_nmax(int n) { return n <= 4 ? 4 : (n <= 8 ? 8 : ((n + 1) & ~1)); }   // the NMAX instantiated for n
static __host__ __device__ inline int fibre2_rb(int n) {
    const int nmax = fibre2_nmax(n);
    return nmax == 10 ? 5 : (nmax == 14 ? 7 : 4);
}
// doubles of shared memory the staged matrices of mode `md` need (0: axis too long for this path)
static __host__ __device__ inline size_t fibre2_smat_doubles(const KronMode &md, int n) {
    if (n > 16 || n < 2 || md.nout != n) return 0;
    const int nmax = fibre2_nmax(n), rb = fibre2_rb(n);
    const int rows = (n + rb - 1) / rb * rb;
    return (size_t)kron_mode_nmats(md) * rows * nmax;
}
#define FIBRE2_SMAT_MAX 2048          // 16 KB of staged matrices at most (SSY (10,)^4: 13 x 100 doubles for all four modes)
// all modes of the view staged side by side (mode m at offset sum of the modes before it): total doubles, 0 when some
// mode does not qualify or the total exceeds the budget - then modes are staged one at a time
static __host__ __device__ inline size_t fibre2_all_doubles(const KronView &kv) {
    size_t tot = 0;
    for (int m = 0; m < kv.n_modes; ++m) {
        const size_t need = fibre2_smat_doubles(kv.modes[m], kv.shape[kv.modes[m].dim]);
        if (need == 0) return 0;
        tot += need;
    }
    return tot <= FIBRE2_SMAT_MAX ? tot : 0;
}


// dynamic shared memory of k_sweep_fused: the resident N-vector + the largest factor matrix the lean
// contraction variants stage (rows 8 IT, pitch 8 IT + 4 with IT = 2, 4, 6, 8; 8 x 8 for the FMA kernel).
// Axes longer than 64 use the cached-load pass, which is not in-place safe: no fused path (returns 0).
static inline int64_t sweep_fused_ldn(int64_t N) { return (N + 1) & ~(int64_t)1; }
static inline size_t sweep_fused_smem(const KronView &kv) {
    size_t smat = 64;
    for (int m = 0; m < kv.n_modes; ++m) {
        const int n = kv.shape[kv.modes[m].dim];
        if (n > KRON_NMAX_LIMIT) return 0;
        if (n >= KRON_TC_MIN) {
            const int it = ((n + 7) / 8 + 1) & ~1;          // even tile count of the lean variants
            const size_t need = (size_t)(8 * it) * (8 * it + 4);
            if (need > smat) smat = need;
        }
    }
    // short axes: all matrices of a mode staged at once for kron_mode_fibre2 (when they fit the budget)
    for (int m = 0; m < kv.n_modes; ++m) {
        const size_t need = fibre2_smat_doubles(kv.modes[m], kv.shape[kv.modes[m].dim]);
        if (need <= FIBRE2_SMAT_MAX && need > smat) smat = need;
    }
    if (fibre2_all_doubles(kv) > smat) smat = fibre2_all_doubles(kv);       // ... of all modes side by side
    return (size_t)(sweep_fused_ldn(kv.N) + smat) * sizeof(double);
}

struct SweepWork {
    double *hl = nullptr, *sc = nullptr, *mz = nullptr;     // per-state base vectors (N)
    double *gamma = nullptr, *theta = nullptr, *beta = nullptr, *last_err = nullptr;   // per column (B)
    unsigned long long *err_bits = nullptr;
    long long *iters = nullptr;
    int *done = nullptr, *n_active = nullptr;
    double *V = nullptr, *Wa = nullptr, *Wb = nullptr;      // panels [B][ldw]
    double *T0 = nullptr, *T1 = nullptr;                    // factor form: mode ping-pong panels
    bool factor_form = false;
    bool fused = false;                                     // factor form with the column resident in shared memory
    KronView kb{};                                          // factor view with a leading column axis
    int64_t ldw = 0;
    cudaStream_t stream = nullptr;
    void free_all() {
        void *ps[] = {hl, sc, mz, gamma, theta, beta, last_err, err_bits, iters, done, n_active, V, Wa, Wb, T0, T1};
        for (void *p : ps) if (p) cudaFreeAsync(p, stream);     // stream-ordered pool (cudaMallocAsync)
        hl = sc = mz = gamma = theta = beta = last_err = V = Wa = Wb = T0 = T1 = nullptr;
        err_bits = nullptr; iters = nullptr; done = n_active = nullptr;
    }
    ~SweepWork() { free_all(); }     // error paths return early: nothing may leak
};

static int sweep_setup(sdfs_op *op, const double *h_prefs, int64_t B, bool panels, SweepWork *w) {
    sdfs_ctx *ctx = op->ctx;
    if (!op->factors || (op->storage != SDFS_STORAGE_DENSE && op->storage != SDFS_STORAGE_KRON))
        return sdfs_set_error(ctx, SDFS_ERR_ARG, "sweep needs an operator built from factors");
    w->stream = ctx->stream;
    w->factor_form = op->storage == SDFS_STORAGE_KRON || op->sweep_form == SDFS_SWEEP_FACTOR;
    if (!w->factor_form && (op->dv.row_begin != 0 || op->dv.row_end != op->dv.N))
        return sdfs_set_error(ctx, SDFS_ERR_ARG, "sweep needs the full P on this rank (columns, not rows, are sharded)");
    const int64_t N = op->kv.N;
    w->ldw = round_up(N, 64);
    const sdfs_factors *f = op->factors;
    CUDA_TRY(ctx, cudaMallocAsync(&w->hl, N * 8, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->sc, N * 8, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->mz, N * 8, ctx->stream));
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
    CUDA_TRY(ctx, cudaMallocAsync(&w->gamma, B * 8, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->theta, B * 8, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->beta, B * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMallocAsync(&w->last_err, B * 8, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->err_bits, B * 8, ctx->stream));
    CUDA_TRY(ctx, cudaMallocAsync(&w->iters, B * 8, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->done, B * 4, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->n_active, 4, ctx->stream));
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
    CUDA_TRY(ctx, cudaMallocAsync(&w->V, pbytes, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(w->V, 0, pbytes, ctx->stream));
    {
        int max_optin = 0;
        CUDA_TRY(ctx, cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        static const bool fused_allowed = !(getenv("SDFS_SWEEP_FUSED") && atoi(getenv("SDFS_SWEEP_FUSED")) == 0);
        const size_t fs = sweep_fused_smem(op->kv);
        w->fused = w->factor_form && fused_allowed && fs > 0 && fs + 1024 <= (size_t)max_optin;
    }
    if (w->factor_form && !w->fused) {
        CUDA_TRY(ctx, cudaMallocAsync(&w->T0, pbytes, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->T1, pbytes, ctx->stream));
        // every mode gains the column index as its outermost free axis (stride ldw): all B columns
        // go through one launch per mode
        w->kb = op->kv;
        for (int m = 0; m < w->kb.n_modes; ++m) {
            KronMode &md = w->kb.modes[m];
            if (md.nF + md.nM >= SDFS_MAX_DIMS)
                return sdfs_set_error(ctx, SDFS_ERR_UNSUPPORTED, "factor-form sweep: too many axes");
            for (int a = md.nF; a > 0; --a) { md.Fshape[a] = md.Fshape[a - 1]; md.Fstride[a] = md.Fstride[a - 1]; }
            md.Fshape[0] = (int)B; md.Fstride[0] = w->ldw;
            md.nF += 1;
            md.Fcount *= B;
        }
    }
    if (panels) {
        CUDA_TRY(ctx, cudaMallocAsync(&w->Wa, pbytes, ctx->stream)); CUDA_TRY(ctx, cudaMallocAsync(&w->Wb, pbytes, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(w->Wa, 0, pbytes, ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(w->Wb, 0, pbytes, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));    // host vectors g/th/be go out of scope
    return SDFS_OK;
}

static int sweep_gemm(sdfs_op *op, SweepWork &w, int64_t B, const double *V, const SweepEpi &ep, const SweepCols &sc) {
    sdfs_ctx *ctx = op->ctx;
    const int64_t N = op->kv.N;
    const int tiles_m = (int)((N + GM - 1) / GM);
    // column-tile width: 128 unless 64 fills the last wave of CTAs markedly better (one CTA per SM;
    // a 64-wide tile does half the work of a 128-wide one at ~90 % of its per-tile efficiency)
    auto wave_eff = [&](int tn) {
        const double tiles = (double)tiles_m * (double)((B + tn - 1) / tn);
        return tiles / (ceil(tiles / ctx->sm_count) * ctx->sm_count);
    };
    const int tn = (0.9 * wave_eff(64) > wave_eff(128)) ? 64 : 128;
    const int tiles_b = (int)((B + tn - 1) / tn);
    const size_t smem = (size_t)GSTAGES * (GM + tn) * GS * sizeof(double);
    const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
    if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    if (tn == 128) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep_gemm<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sweep_gemm<128><<<tiles_m * tiles_b, GTHREADS, smem, ctx->stream>>>(op->dv.P, N, op->dv.ld, V, B, w.ldw, w.sc, w.mz, ep, sc, tiles_b);
    } else {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep_gemm<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sweep_gemm<64><<<tiles_m * tiles_b, GTHREADS, smem, ctx->stream>>>(op->dv.P, N, op->dv.ld, V, B, w.ldw, w.sc, w.mz, ep, sc, tiles_b);
    }
    if (prof) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

// Factor form: the GEMM epilogue as an elementwise kernel.  One CTA per (column, chunk of rows), so the
// per-column scalars are uniform in the CTA and the sup-norm needs one atomic per CTA.
#define SWE_ROWS 2048
__global__ void __launch_bounds__(256) k_sweep_epi_ew(int64_t N, int64_t ldw, const double *__restrict__ S,
                                                       const double *__restrict__ sig_c, const double *__restrict__ mz,
                                                       SweepEpi ep, SweepCols sc) {
    const int64_t chunks = (N + SWE_ROWS - 1) / SWE_ROWS;
    const int64_t b = blockIdx.x / chunks, n0 = (blockIdx.x % chunks) * SWE_ROWS;
    const int64_t n1 = n0 + SWE_ROWS < N ? n0 + SWE_ROWS : N;
    const double *Sb = S + b * ldw;
    if (ep.mode == 2) {
        for (int64_t n = n0 + threadIdx.x; n < n1; n += blockDim.x)
            ep.out0[b * ldw + n] = ep.D[b * ldw + n] * Sb[n] - ep.Vsub[b * ldw + n];
        return;
    }
    const double g = sc.gamma[b], th = sc.theta[b], be = sc.beta[b];
    const int frozen = sc.done ? sc.done[b] : 0;
    const double omg = 1.0 - g, inv_th = 1.0 / th;
    double emax = 0.0;
    for (int64_t n = n0 + threadIdx.x; n < n1; n += blockDim.x) {
        // log-domain form of the GEMM epilogue: log(a_row S) = log S + (1/2)((1-gamma) sigma_c)^2 + (1-gamma)(mu_c + z),
        // so T needs one log and one exp per element instead of two exp and a pow
        const double t2 = omg * sig_c[n];
        const double la = 0.5 * (t2 * t2) + omg * mz[n];        // log a_row
        const double w_old = ep.W[b * ldw + n];
        const double ls = log(Sb[n]) + la;                      // log(a_row S); NaN for S < 0 as pow would give
        double y = 1.0 + be * exp(inv_th * ls);
        if (ep.mode == 0) {
            if (frozen) y = w_old;
            ep.out0[b * ldw + n] = y;
            const double d = fabs(y - w_old);
            emax = (d != d || emax != emax) ? d + emax : fmax(emax, d);   // NaN propagates
        } else {
            ep.out0[b * ldw + n] = y - w_old;
            ep.out1[b * ldw + n] = be * exp((1.0 - th) * inv_th * ls + la);
        }
    }
    if (ep.mode == 0 && ep.err_bits) {
        __shared__ double red[8];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, emax, o);
            emax = (other != other || emax != emax) ? other + emax : fmax(emax, other);
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = emax;
        __syncthreads();
        if (threadIdx.x == 0) {
            double e = red[0];
            for (int i = 1; i < (int)(blockDim.x >> 5); ++i) e = (red[i] != red[i] || e != e) ? red[i] + e : fmax(e, red[i]);
            atomicMax(ep.err_bits + b, (unsigned long long)__double_as_longlong(fabs(e)));
        }
    }
}

int launch_kron_mode(sdfs_ctx *ctx, const KronView &kv, int m, const double *in, double *out);   // ops.cu

// S = P V for all columns through the Markov factors (one launch per mode, columns batched), then the epilogue
static int sweep_factor(sdfs_op *op, SweepWork &w, int64_t B, const double *V, const SweepEpi &ep, const SweepCols &sc) {
    sdfs_ctx *ctx = op->ctx;
    const int64_t N = op->kv.N;
    const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
    if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    const double *in = V;
    for (int m = 0; m < w.kb.n_modes; ++m) {
        double *out = (m & 1) ? w.T1 : w.T0;
        TRY(launch_kron_mode(ctx, w.kb, m, in, out));
        in = out;
    }
    const int64_t chunks = (N + SWE_ROWS - 1) / SWE_ROWS;
    k_sweep_epi_ew<<<(unsigned)(B * chunks), 256, 0, ctx->stream>>>(N, w.ldw, in, w.sc, w.mz, ep, sc);
    if (prof) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

// ---------------------------------------------------------------------------
// Fused factor-form application for grids whose state vector fits in shared memory (BASELINE config 5:
// N = 10 000 -> 80 KB).  One CTA per parameter column keeps the column resident: (optional prologue) ->
// every mode contraction IN PLACE in shared memory (a contraction maps each fibre onto itself and a
// tile loads all its inputs before it stores, so no second buffer is needed; same tensor-core tiles as
// k_kron_mode, the CTA taking the whole tile range) -> epilogue.  A column is read once and written
// once per application instead of once per mode; two CTAs per SM overlap one column's global loads with
// the other's contractions.  Columns that have converged (SA) cost one copy.
// ---------------------------------------------------------------------------
// Short axes (n <= 16) on a vector resident in shared memory: register-blocked FMA contraction, two fibres per thread.
// A thread loads both fibres into registers (so the contraction is in place: nobody else touches these fibres), then
// forms RB output rows at a time: every matrix element fetched (LDS.128, the same address across the warp while its
// fibres share a matrix) feeds two FMAs per fibre - 2 x NMAX x RB FMAs per NMAX x RB / 2 shared-memory loads, which puts
// the loop on the fp64 pipe instead of the shared-memory pipe (kron_mode_fibre: one fibre per thread, rows padded to 4 and
// columns to 16, one matrix per work item with two block barriers each - 14 clocks per fibre per SM at n = 10).
// ALL matrices of the mode are staged once (row pitch NMAX, rows padded to a multiple of RB with zeros), so the
// whole mode is one pass over the fibres without block barriers inside.
// Fibre pairs: thread t takes fibres q and q + ceil(F / 2) of one matrix combination, so neighbouring threads read
// neighbouring fibres (conflict-free for every mode but the innermost axis, 4-way there).
// stage every matrix of the mode: row pitch NMAX, rows padded to a multiple of RB, zeros outside n x n
__device__ __forceinline__ void fibre2_stage(const KronMode &md, int n, double *smat) {
    const int NMAX = fibre2_nmax(n), RB = fibre2_rb(n);
    const int rows = (n + RB - 1) / RB * RB;
    const int msz = rows * NMAX;
    const int nmats = kron_mode_nmats(md);
    for (int e = threadIdx.x; e < nmats * msz; e += blockDim.x) {
        const int mt = e / msz, r = e - mt * msz, i = r / NMAX, j = r - i * NMAX;
        double v = 0.0;
        if (i < n && j < n) {
            v = md.mat[((long long)mt * n + i) * n + j];
            if (md.colscale) v *= md.colscale[j];
        }
        smat[e] = v;
    }
}

template <int NMAX, int RB>
__device__ __forceinline__ void kron_mode_fibre2(const KronMode &md, int n, double *cur, const double *smat) {
    const int rows = (n + RB - 1) / RB * RB;
    const int msz = rows * NMAX;
    const unsigned F = (unsigned)md.Fcount, half = (F + 1) >> 1;
    const unsigned pairs = (unsigned)md.Mcount * half;
    const int stride = (int)md.stride;            // the vector lives in shared memory: 32-bit element indices
    for (unsigned q = threadIdx.x; q < pairs; q += blockDim.x) {
        const unsigned mc = q / half, f0 = q - mc * half, f1 = f0 + half;
        unsigned rem = mc;
        int mbase = (int)md.base_off;
        int mat = 0;
        for (int a = md.nM - 1; a >= 0; --a) {
            const unsigned sh = (unsigned)md.Mshape[a], c = rem % sh;
            rem /= sh;
            mat += (int)c * md.Mmat[a];
            mbase += (int)c * (int)md.Mstride[a];
        }
        int b0 = mbase, b1 = mbase;
        unsigned r0 = f0, r1 = f1 < F ? f1 : f0;
        for (int a = md.nF - 1; a >= 0; --a) {
            const unsigned sh = (unsigned)md.Fshape[a];
            const unsigned c0 = r0 % sh, c1 = r1 % sh;
            r0 /= sh; r1 /= sh;
            b0 += (int)c0 * (int)md.Fstride[a];
            b1 += (int)c1 * (int)md.Fstride[a];
        }
        const bool two = f1 < F;
        double x0[NMAX], x1[NMAX];
        double *p0 = cur + b0, *p1 = cur + b1;
#pragma unroll
        for (int j = 0; j < NMAX; ++j) {
            x0[j] = (j < n) ? p0[j * stride] : 0.0;
            x1[j] = (j < n) ? p1[j * stride] : 0.0;
        }
        const double2 *mrow = reinterpret_cast<const double2 *>(smat + mat * msz);
        for (int rb = 0; rb < rows; rb += RB) {
            double a0[RB], a1[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) a0[r] = a1[r] = 0.0;
#pragma unroll
            for (int j = 0; j < NMAX / 2; ++j) {
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const double2 mm = mrow[(rb + r) * (NMAX / 2) + j];
                    a0[r] = fma(mm.x, x0[2 * j], a0[r]); a1[r] = fma(mm.x, x1[2 * j], a1[r]);
                    a0[r] = fma(mm.y, x0[2 * j + 1], a0[r]); a1[r] = fma(mm.y, x1[2 * j + 1], a1[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (rb + r < n) {
                    p0[(rb + r) * stride] = a0[r];
                    if (two) p1[(rb + r) * stride] = a1[r];
                }
        }
    }
}

// one mode of the resident vector on the register-blocked FMA path (matrices already staged at `smat`)
__device__ __forceinline__ void fibre2_dispatch(const KronMode &md, int n, double *cur, const double *smat) {
    switch (fibre2_nmax(n)) {
    case 4: kron_mode_fibre2<4, 4>(md, n, cur, smat); break;
    case 8: kron_mode_fibre2<8, 4>(md, n, cur, smat); break;
    case 10: kron_mode_fibre2<10, 5>(md, n, cur, smat); break;
    case 12: kron_mode_fibre2<12, 4>(md, n, cur, smat); break;
    case 14: kron_mode_fibre2<14, 7>(md, n, cur, smat); break;
    default: kron_mode_fibre2<16, 4>(md, n, cur, smat); break;
    }
}

#define SWF_THREADS 256
#define SWF_UNROLL 8            // independent global loads in flight per thread in the load / epilogue loops
__global__ void __launch_bounds__(SWF_THREADS, 2)
k_sweep_fused(const __grid_constant__ KronView kv, int64_t ldw, int64_t ldn, const double *__restrict__ in_panel, int prologue, int fibre2,
              const double *__restrict__ h_lam, const double *__restrict__ sig_c, const double *__restrict__ mz,
              SweepEpi ep, SweepCols sc) {
    extern __shared__ __align__(16) double fsm[];
    const int64_t N = kv.N;
    double *cur = fsm, *smat = fsm + ldn;
    const int64_t b = blockIdx.x;
    const double th = sc.theta[b];
    const int frozen = (ep.mode == 0 && sc.done) ? sc.done[b] : 0;
    const int64_t step = (int64_t)SWF_UNROLL * blockDim.x;
    if (frozen) {                                   // converged column: W' = W, error 0
        for (int64_t n = threadIdx.x; n < N; n += blockDim.x) ep.out0[b * ldw + n] = ep.W[b * ldw + n];
        return;
    }
    const double *col = in_panel + b * ldw;
    // plain column load (every Krylov mat-vec): ONE bulk async copy, all of the column in flight at once, completing on
    // an mbarrier while the threads stage the factor matrices.  The streams the epilogue will read are prefetched into
    // L2 meanwhile (they are consumed ~20 us later, after the contractions).
    __shared__ uint64_t col_bar;
    const uint32_t col_bytes = (uint32_t)(ldn * sizeof(double));          // even element count; the panel pitch covers it
    const bool bulk = !prologue && (((uintptr_t)col | (uintptr_t)cur) & 15) == 0 && ldn <= ldw;
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&col_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&col_bar, col_bytes);
            bulk_g2s(cur, col, col_bytes, &col_bar);
            const uint32_t pf = (uint32_t)((N * sizeof(double)) & ~(size_t)15);
            if (pf) {
                if (ep.mode == 2) {
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ep.D + b * ldw), "r"(pf) : "memory");
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ep.Vsub + b * ldw), "r"(pf) : "memory");
                } else {
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ep.W + b * ldw), "r"(pf) : "memory");
                }
            }
        }
    } else {
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double v[SWF_UNROLL], h[SWF_UNROLL];
#pragma unroll
            for (int u = 0; u < SWF_UNROLL; ++u) {
                const int64_t n = n0 + (int64_t)u * blockDim.x;
                v[u] = n < N ? col[n] : 1.0;
                h[u] = (prologue && n < N) ? h_lam[n] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < SWF_UNROLL; ++u) {
                const int64_t n = n0 + (int64_t)u * blockDim.x;
                if (n < N) cur[n] = prologue ? pow_pos_off(v[u], th, h[u]) : v[u];    // exp(theta h) w^theta
            }
        }
    }
    // short axes throughout: the matrices of all modes are staged now, next to the column load, and every mode is
    // one pass over its fibres with a single block barrier behind it
    const bool all_staged = fibre2 && fibre2_all_doubles(kv) > 0;
    if (all_staged) {
        double *sm = smat;
        for (int m = 0; m < kv.n_modes; ++m) {
            const int n = kv.shape[kv.modes[m].dim];
            fibre2_stage(kv.modes[m], n, sm);
            sm += fibre2_smat_doubles(kv.modes[m], n);
        }
    }
    __syncthreads();                    // staged matrices (and, for thread 0's barrier init, the barrier itself) visible
    if (bulk) mbar_wait(&col_bar, 0);   // the column has landed
    {
        const double *sm = smat;
        for (int m = 0; m < kv.n_modes; ++m) {
            const KronMode &md = kv.modes[m];
            const int n = kv.shape[md.dim];
            const size_t need = fibre2 ? fibre2_smat_doubles(md, n) : 0;
            if (all_staged) {
                fibre2_dispatch(md, n, cur, sm);
                sm += need;
            } else if (need > 0 && need <= FIBRE2_SMAT_MAX) {
                fibre2_stage(md, n, smat);
                __syncthreads();
                fibre2_dispatch(md, n, cur, smat);
            } else {
                kron_mode_apply<false, true>(kv, m, cur, smat, [&](int64_t idx, double s) { cur[idx] = s; }, KronShare(0, 1));
            }
            __syncthreads();
        }
    }
    // epilogue on the resident contraction (same arithmetic as k_sweep_epi_ew)
    if (ep.mode == 2) {
        const double *Db = ep.D + b * ldw, *Vb = ep.Vsub + b * ldw;
        double *Ob = ep.out0 + b * ldw;
        if (((((uintptr_t)Db) | ((uintptr_t)Vb) | ((uintptr_t)Ob)) & 15) == 0) {
            // 16-byte accesses, SWF_UNROLL / 2 x 2 loads of each stream in flight per thread
            const int64_t N2 = N >> 1;
            const double2 *D2 = reinterpret_cast<const double2 *>(Db), *V2 = reinterpret_cast<const double2 *>(Vb);
            const double2 *C2 = reinterpret_cast<const double2 *>(cur);
            double2 *O2 = reinterpret_cast<double2 *>(Ob);
            constexpr int U2 = SWF_UNROLL / 2;
            for (int64_t n0 = threadIdx.x; n0 < N2; n0 += (int64_t)U2 * blockDim.x) {
                double2 d[U2], vs[U2];
#pragma unroll
                for (int u = 0; u < U2; ++u) {
                    const int64_t n = n0 + (int64_t)u * blockDim.x;
                    d[u] = n < N2 ? D2[n] : make_double2(0.0, 0.0);
                    vs[u] = n < N2 ? V2[n] : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < U2; ++u) {
                    const int64_t n = n0 + (int64_t)u * blockDim.x;
                    if (n < N2) {
                        const double2 c = C2[n];
                        O2[n] = make_double2(d[u].x * c.x - vs[u].x, d[u].y * c.y - vs[u].y);
                    }
                }
            }
            if ((N & 1) && threadIdx.x == 0) Ob[N - 1] = Db[N - 1] * cur[N - 1] - Vb[N - 1];
            return;
        }
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double d[SWF_UNROLL], vs[SWF_UNROLL];
#pragma unroll
            for (int u = 0; u < SWF_UNROLL; ++u) {
                const int64_t n = n0 + (int64_t)u * blockDim.x;
                d[u] = n < N ? Db[n] : 0.0;
                vs[u] = n < N ? Vb[n] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < SWF_UNROLL; ++u) {
                const int64_t n = n0 + (int64_t)u * blockDim.x;
                if (n < N) Ob[n] = d[u] * cur[n] - vs[u];
            }
        }
        return;
    }
    const double g = sc.gamma[b], be = sc.beta[b];
    const double omg = 1.0 - g, inv_th = 1.0 / th;
    double emax = 0.0;
    for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
        double wo[SWF_UNROLL], sg[SWF_UNROLL], zz[SWF_UNROLL];
#pragma unroll
        for (int u = 0; u < SWF_UNROLL; ++u) {
            const int64_t n = n0 + (int64_t)u * blockDim.x;
            wo[u] = n < N ? ep.W[b * ldw + n] : 0.0;
            sg[u] = n < N ? sig_c[n] : 0.0;
            zz[u] = n < N ? mz[n] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < SWF_UNROLL; ++u) {
            const int64_t n = n0 + (int64_t)u * blockDim.x;
            if (n >= N) continue;
            const double t2 = omg * sg[u];
            const double la = 0.5 * (t2 * t2) + omg * zz[u];
            if (ep.mode == 0) {
                // SA step: the table-driven exp(e (log x + off)) (common.cuh), as in k_sweep_sa_col
                const double y = 1.0 + be * pow_pos_off(cur[n], inv_th, la);
                ep.out0[b * ldw + n] = y;
                const double dd = fabs(y - wo[u]);
                emax = (dd != dd || emax != emax) ? dd + emax : fmax(emax, dd);   // NaN propagates
            } else {
                const double ls = log(cur[n]) + la;
                const double y = 1.0 + be * exp(inv_th * ls);
                ep.out0[b * ldw + n] = y - wo[u];
                ep.out1[b * ldw + n] = be * exp((1.0 - th) * inv_th * ls + la);
            }
        }
    }
    if (ep.mode == 0 && ep.err_bits) {
        __shared__ double red[SWF_THREADS / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, emax, o);
            emax = (other != other || emax != emax) ? other + emax : fmax(emax, other);
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = emax;
        __syncthreads();
        if (threadIdx.x == 0) {
            double e = red[0];
            for (int i = 1; i < (int)(blockDim.x >> 5); ++i) e = (red[i] != red[i] || e != e) ? red[i] + e : fmax(e, red[i]);
            atomicMax(ep.err_bits + b, (unsigned long long)__double_as_longlong(fabs(e)));
        }
    }
}

static int sweep_fused(sdfs_op *op, SweepWork &w, int64_t B, const double *in_panel, int prologue, const SweepEpi &ep,
                       const SweepCols &sc) {
    sdfs_ctx *ctx = op->ctx;
    const size_t smem = sweep_fused_smem(op->kv);
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
    if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    // SDFS_SWEEP_FIBRE2=0: the round-1 thread-per-fibre contraction for short axes (A/B switch)
    static const int fibre2 = !(getenv("SDFS_SWEEP_FIBRE2") && atoi(getenv("SDFS_SWEEP_FIBRE2")) == 0);
    k_sweep_fused<<<(unsigned)B, SWF_THREADS, smem, ctx->stream>>>(op->kv, w.ldw, sweep_fused_ldn(op->kv.N), in_panel, prologue, fibre2, w.hl, w.sc,
                                                                  w.mz, ep, sc);
    if (prof) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

static inline int sweep_PV(sdfs_op *op, SweepWork &w, int64_t B, const double *V, const SweepEpi &ep, const SweepCols &sc) {
    if (w.fused) return sweep_fused(op, w, B, V, 0, ep, sc);
    return w.factor_form ? sweep_factor(op, w, B, V, ep, sc) : sweep_gemm(op, w, B, V, ep, sc);
}

static inline int panel_grid(sdfs_ctx *ctx, int64_t tot) {
    const int64_t g = (tot + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

static int sweep_step(sdfs_op *op, SweepWork &w, int64_t B, const double *Win, double *Wout, bool track) {
    sdfs_ctx *ctx = op->ctx;
    const int64_t N = op->kv.N;
    SweepCols sc{w.gamma, w.theta, w.beta, track ? w.done : nullptr};
    SweepEpi ep{0, Win, Wout, nullptr, nullptr, nullptr, track ? w.err_bits : nullptr};
    if (w.fused) return sweep_fused(op, w, B, Win, 1, ep, sc);         // prologue fused: reads W itself
    k_sweep_prologue<<<panel_grid(ctx, N * B), 256, 0, ctx->stream>>>(N, B, w.ldw, w.hl, Win, sc, w.V);
    ctx->launches++;
    return sweep_PV(op, w, B, w.V, ep, sc);
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

// successive approximation of every column inside one CTA (defined next to the Newton sweep's fused inner solve)
static bool sweep_sa_cols_applicable(const sdfs_op *op, const SweepWork &w);
static int sweep_sa_cols(sdfs_op *op, SweepWork &w, int64_t B, double *W, double tol, long long max_iter);

extern "C" {

int sdfs_sweep_set_form(sdfs_op *op, int form) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_sweep_set_form: NULL op");
    ARG_CHECK(op->ctx, form == SDFS_SWEEP_DENSE || form == SDFS_SWEEP_FACTOR);
    if (form == SDFS_SWEEP_DENSE && op->storage != SDFS_STORAGE_DENSE)
        return sdfs_set_error(op->ctx, SDFS_ERR_ARG, "the dense (GEMM) sweep needs a dense operator");
    op->sweep_form = form;
    return SDFS_OK;
}

int sdfs_sweep_apply_T(sdfs_op *op, const double *h_prefs, int64_t B, const double *d_W_in, double *d_W_out) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_sweep_apply_T: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, h_prefs && d_W_in && d_W_out && B >= 1);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    SweepWork w;
    int rc = sweep_setup(op, h_prefs, B, true, &w);
    const int64_t N = op->kv.N;
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
    const int64_t N = op->kv.N;
    const int grid = ctx->sm_count * 8;
    double *cur = w.Wa, *nxt = w.Wb;
    if (rc == SDFS_OK) {
        k_fill_panel<<<grid, 256, 0, ctx->stream>>>(cur, N, B, w.ldw, w_init);
        ctx->launches++;
        int active = (int)B;
        int64_t steps = 0;
        if (max_iter > 0 && sweep_sa_cols_applicable(op, w)) {
            rc = sweep_sa_cols(op, w, B, cur, tol, (long long)max_iter);      // in place: `cur` holds the fixed points
            active = 0;
        }
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

// ===========================================================================
// Newton mode of the sweep: every parameter column runs the reference's Newton iteration
// (solvers.py:51-95) with its own BiCGSTAB (JAX recurrence, per-column scalars and stopping
// tests); the columns advance in lockstep so that every Krylov mat-vec of all columns is ONE
// fp64 tensor-core GEMM  (J_g v)[b] = d_b .* P (c_b .* v_b) - v_b.  Columns whose inner or
// outer loop has finished are masked on the device.  One CTA per column handles the vector
// updates and the (deterministic, fixed-order) dot products of that column.
// ===========================================================================
#define SWN_THREADS 256

struct SwnState {                 // per-column device arrays (length B)
    double *rho, *alpha, *omega, *rs, *rho_next, *atol2, *last_err;
    long long *k, *outer_it, *inner_total;
    int *exit_early, *in_active, *out_active;
    int *n_in_active, *n_out_active;           // global counters
};
struct SwnPanels {                // [B][ldw] each
    double *W, *G, *C, *D, *X, *R, *Rh, *Pv, *Q, *Sv, *T, *Xin;
};

__device__ __forceinline__ double cta_reduce_sum(double v, double *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < SWN_THREADS / 32; ++w) s += sm[w];
    return s;
}
__device__ __forceinline__ double cta_reduce_nanmax(double v, double *sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_nanmax(v);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < SWN_THREADS / 32; ++w) s = nanmax(s, sm[w]);
    return s;
}

// Xin = a_col_b .* W^theta_b ; C = a_col_b .* W^(theta_b - 1)   (columns still iterating)
__global__ void k_swn_prologue(int64_t N, int64_t ldw, const double *__restrict__ h_lam, SweepCols sc, SwnState st, SwnPanels p) {
    const int64_t b = blockIdx.x;
    if (!st.out_active[b]) return;
    const double th = sc.theta[b];
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) {
        const double w = p.W[b * ldw + n], ac = exp(th * h_lam[n]);
        p.Xin[b * ldw + n] = ac * pow(w, th);
        p.C[b * ldw + n] = ac * pow(w, th - 1.0);
    }
}

// Krylov start: r = rhat = p = q = g, x = 0, <g,g>, stopping threshold, scalars
__global__ void k_swn_init(int64_t N, int64_t ldw, double rtol, double atol, SwnState st, SwnPanels p) {
    __shared__ double sm[SWN_THREADS / 32];
    const int64_t b = blockIdx.x;
    if (!st.out_active[b]) { if (threadIdx.x == 0) st.in_active[b] = 0; return; }
    double acc = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) {
        const double g = p.G[b * ldw + n];
        p.R[b * ldw + n] = g; p.Rh[b * ldw + n] = g; p.Pv[b * ldw + n] = g; p.Q[b * ldw + n] = g;
        p.X[b * ldw + n] = 0.0;
        acc += g * g;
    }
    const double bs = cta_reduce_sum(acc, sm);
    if (threadIdx.x == 0) {
        const double a2 = fmax(rtol * rtol * bs, atol * atol);
        st.atol2[b] = a2; st.rs[b] = bs; st.rho_next[b] = bs;
        st.rho[b] = 1.0; st.alpha[b] = 1.0; st.omega[b] = 1.0; st.k[b] = 0;
        const int act = (bs > a2) ? 1 : 0;          // NaN -> inactive (loop exits, like the reference)
        st.in_active[b] = act;
        if (act) atomicAdd(st.n_in_active, 1);
    }
}

// p = r + beta (p - omega q) ; Xin = c .* p
__global__ void k_swn_phase1(int64_t N, int64_t ldw, SwnState st, SwnPanels p) {
    const int64_t b = blockIdx.x;
    if (!st.in_active[b]) return;
    const double rho_ = st.rho_next[b];
    const double beta = rho_ / st.rho[b] * st.alpha[b] / st.omega[b];
    const double omega = st.omega[b];
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) {
        const int64_t e = b * ldw + n;
        const double pn = p.R[e] + beta * (p.Pv[e] - omega * p.Q[e]);
        p.Pv[e] = pn;
        p.Xin[e] = p.C[e] * pn;
    }
}

// alpha = rho'/<rhat,q> ; s = r - alpha q ; <s,s> ; Xin = c .* s
__global__ void k_swn_phase3(int64_t N, int64_t ldw, SwnState st, SwnPanels p) {
    __shared__ double sm[SWN_THREADS / 32];
    const int64_t b = blockIdx.x;
    if (!st.in_active[b]) return;
    double acc = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) acc += p.Rh[b * ldw + n] * p.Q[b * ldw + n];
    const double rq = cta_reduce_sum(acc, sm);
    const double alpha_ = st.rho_next[b] / rq;
    double ss = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) {
        const int64_t e = b * ldw + n;
        const double sn = p.R[e] - alpha_ * p.Q[e];
        p.Sv[e] = sn;
        p.Xin[e] = p.C[e] * sn;
        ss += sn * sn;
    }
    ss = cta_reduce_sum(ss, sm);
    if (threadIdx.x == 0) {
        st.alpha[b] = alpha_;                       // rho (old) is still needed? no: beta was formed in phase 1
        st.exit_early[b] = (ss < st.atol2[b]) ? 1 : 0;
    }
}

// omega = <t,s>/<t,t> ; x, r updates ; <r,r>, <rhat,r> ; k and the loop condition
__global__ void k_swn_phase5(int64_t N, int64_t ldw, long long maxiter, SwnState st, SwnPanels p) {
    __shared__ double sm[SWN_THREADS / 32];
    const int64_t b = blockIdx.x;
    if (!st.in_active[b]) return;
    double ts = 0.0, tt = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) {
        const int64_t e = b * ldw + n;
        const double tn = p.T[e];
        ts += tn * p.Sv[e];
        tt += tn * tn;
    }
    ts = cta_reduce_sum(ts, sm);
    tt = cta_reduce_sum(tt, sm);
    const double omega_ = ts / tt, alpha_ = st.alpha[b];
    const int early = st.exit_early[b];
    double rr = 0.0, rhr = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) {
        const int64_t e = b * ldw + n;
        const double pn = p.Pv[e], sn = p.Sv[e];
        double xn, rn;
        if (early) { xn = p.X[e] + alpha_ * pn; rn = sn; }
        else { xn = p.X[e] + (alpha_ * pn + omega_ * sn); rn = sn - omega_ * p.T[e]; }
        p.X[e] = xn;
        p.R[e] = rn;
        rr += rn * rn;
        rhr += p.Rh[e] * rn;
    }
    rr = cta_reduce_sum(rr, sm);
    rhr = cta_reduce_sum(rhr, sm);
    if (threadIdx.x == 0) {
        const double rho_ = st.rho_next[b];
        long long k_ = (omega_ == 0.0 || alpha_ == 0.0) ? -11 : st.k[b] + 1;
        if (rho_ == 0.0) k_ = -10;
        st.k[b] = k_;
        st.omega[b] = omega_;
        st.rho[b] = rho_;
        st.rs[b] = rr;
        st.rho_next[b] = rhr;
        if (!(rr > st.atol2[b] && k_ < maxiter && k_ >= 0)) {
            st.in_active[b] = 0;
            st.inner_total[b] += (k_ > 0 ? k_ : 0);
            atomicSub(st.n_in_active, 1);
        }
    }
}

// ---------------------------------------------------------------------------
// The whole inner BiCGSTAB solve of one column in ONE CTA (grids whose column fits in shared memory and whose axes
// are all short: the k_sweep_fused / kron_mode_fibre2 case).  A column's Krylov iteration needs nothing from the
// other columns, and the CTA that contracts a column already holds all of it, so phases 1, 3 and 5 above run as
// loops of the same CTA around the two resident contractions, their dot products as block reductions: no launch
// boundaries, no host polls, the factor matrices staged once per solve, and the mat-vec input / the vectors q and t
// never travel through global memory (q is stored once for the next iteration's p update, t not at all).
// Same arithmetic per element and the same per-thread accumulation order as k_swn_phase1/3/5 + k_sweep_fused
// (256 threads, element n on thread n mod 256, partials in n order), so every scalar - and with it every
// column's iteration count - is bit-identical to the phase-kernel path.
// ---------------------------------------------------------------------------
#define SWI_U 4
__global__ void __launch_bounds__(SWF_THREADS, 2)
k_swn_inner_fused(const __grid_constant__ KronView kv, int64_t ldw, int64_t ldn, long long maxiter, SwnState st, SwnPanels p,
                  int *max_k) {
    extern __shared__ __align__(16) double fsm[];
    __shared__ double sm[SWN_THREADS / 32];
    static_assert(SWF_THREADS == SWN_THREADS, "the column loops assume the phase kernels' thread count");
    const int64_t N = kv.N;
    const int64_t b = blockIdx.x;
    if (!st.in_active[b]) return;
    double *cur = fsm, *smat = fsm + ldn;
    {
        double *sm_m = smat;
        for (int m = 0; m < kv.n_modes; ++m) {
            const int n = kv.shape[kv.modes[m].dim];
            fibre2_stage(kv.modes[m], n, sm_m);
            sm_m += fibre2_smat_doubles(kv.modes[m], n);
        }
    }
    const double *R = p.R + b * ldw, *Rh = p.Rh + b * ldw, *C = p.C + b * ldw, *D = p.D + b * ldw;
    double *Rw = p.R + b * ldw, *Pv = p.Pv + b * ldw, *Q = p.Q + b * ldw, *Sv = p.Sv + b * ldw, *X = p.X + b * ldw;
    double rho = st.rho[b], alpha = st.alpha[b], omega = st.omega[b], rho_next = st.rho_next[b], rs = st.rs[b];
    const double atol2 = st.atol2[b];
    long long k = st.k[b];
    const int64_t step = (int64_t)SWI_U * SWN_THREADS;
    auto contract = [&]() {
        const double *sm_m = smat;
        __syncthreads();
        for (int m = 0; m < kv.n_modes; ++m) {
            const KronMode &md = kv.modes[m];
            const int n = kv.shape[md.dim];
            fibre2_dispatch(md, n, cur, sm_m);
            sm_m += fibre2_smat_doubles(md, n);
            __syncthreads();
        }
    };
    for (;;) {
        // ---- phase 1: p = r + beta (p - omega q); mat-vec input c .* p
        const double beta = rho_next / rho * alpha / omega;
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double r[SWI_U], pv[SWI_U], q[SWI_U], c[SWI_U];
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) { r[u] = R[n]; pv[u] = Pv[n]; q[u] = Q[n]; c[u] = C[n]; }
            }
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) {
                    const double pn = r[u] + beta * (pv[u] - omega * q[u]);
                    Pv[n] = pn;
                    cur[n] = c[u] * pn;
                }
            }
        }
        contract();
        // ---- q = d .* S - p (the Krylov epilogue of k_sweep_fused), <rhat, q>
        double acc = 0.0;
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double d[SWI_U], pv[SWI_U], rh[SWI_U];
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) { d[u] = D[n]; pv[u] = Pv[n]; rh[u] = Rh[n]; }
            }
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) {
                    const double qn = d[u] * cur[n] - pv[u];
                    Q[n] = qn;
                    cur[n] = qn;
                    acc += rh[u] * qn;
                }
            }
        }
        const double rq = cta_reduce_sum(acc, sm);
        // ---- phase 3: alpha, s = r - alpha q, <s, s>, mat-vec input c .* s
        const double alpha_ = rho_next / rq;
        double ss = 0.0;
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double r[SWI_U], c[SWI_U];
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) { r[u] = R[n]; c[u] = C[n]; }
            }
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) {
                    const double sn = r[u] - alpha_ * cur[n];
                    Sv[n] = sn;
                    cur[n] = c[u] * sn;
                    ss += sn * sn;
                }
            }
        }
        ss = cta_reduce_sum(ss, sm);
        const int early = (ss < atol2) ? 1 : 0;
        contract();
        // ---- t = d .* S - s (kept in shared memory only), <t, s>, <t, t>
        double ts = 0.0, tt = 0.0;
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double d[SWI_U], sv[SWI_U];
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) { d[u] = D[n]; sv[u] = Sv[n]; }
            }
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) {
                    const double tn = d[u] * cur[n] - sv[u];
                    cur[n] = tn;
                    ts += tn * sv[u];
                    tt += tn * tn;
                }
            }
        }
        ts = cta_reduce_sum(ts, sm);
        tt = cta_reduce_sum(tt, sm);
        // ---- phase 5: omega, x and r updates, <r, r>, <rhat, r>, loop condition
        const double omega_ = ts / tt;
        double rr = 0.0, rhr = 0.0;
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double pv[SWI_U], sv[SWI_U], x[SWI_U], rh[SWI_U];
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) { pv[u] = Pv[n]; sv[u] = Sv[n]; x[u] = X[n]; rh[u] = Rh[n]; }
            }
#pragma unroll
            for (int u = 0; u < SWI_U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWN_THREADS;
                if (n < N) {
                    double xn, rn;
                    if (early) { xn = x[u] + alpha_ * pv[u]; rn = sv[u]; }
                    else { xn = x[u] + (alpha_ * pv[u] + omega_ * sv[u]); rn = sv[u] - omega_ * cur[n]; }
                    X[n] = xn;
                    Rw[n] = rn;
                    rr += rn * rn;
                    rhr += rh[u] * rn;
                }
            }
        }
        rr = cta_reduce_sum(rr, sm);
        rhr = cta_reduce_sum(rhr, sm);
        const double rho_ = rho_next;
        long long k_ = (omega_ == 0.0 || alpha_ == 0.0) ? -11 : k + 1;
        if (rho_ == 0.0) k_ = -10;
        k = k_; omega = omega_; alpha = alpha_; rho = rho_; rs = rr; rho_next = rhr;
        if (!(rr > atol2 && k_ < maxiter && k_ >= 0)) break;
        __syncthreads();            // this iteration's stores to r, p, q are read by other threads' elements? no: same thread per
                                    // element throughout; the barrier only keeps `cur` reuse ordered for the next phase 1
    }
    if (threadIdx.x == 0) {
        st.k[b] = k; st.omega[b] = omega; st.alpha[b] = alpha; st.rho[b] = rho; st.rs[b] = rs; st.rho_next[b] = rho_next;
        st.in_active[b] = 0;
        st.inner_total[b] += (k > 0 ? k : 0);
        atomicSub(st.n_in_active, 1);
        atomicMax(max_k, (int)(k > 0 ? k : 1));
    }
}

// ---------------------------------------------------------------------------
// Successive approximation of one column in ONE CTA, all iterations (solvers.py:19-48 per column): the same case as
// k_swn_inner_fused (column in shared memory, short axes).  Per step: the resident contractions, then one loop that
// evaluates the epilogue y = 1 + beta exp((log S + log a_row) / theta), the error against the stored w, stores y and
// writes the NEXT step's contraction input exp(theta (h_lambda + log y)) straight back into shared memory - the
// arithmetic of k_sweep_fused's epilogue and prologue, element by element, so iterates, errors and iteration counts are
// bit-identical to the launch-per-step path.  No launches, no host polls, no panel round trips: a step reads w (its own
// 80 KB, L2-resident) and the three state vectors, and writes w.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SWF_THREADS, 2)
k_sweep_sa_col(const __grid_constant__ KronView kv, int64_t ldw, int64_t ldn, double *__restrict__ W,
               const double *__restrict__ h_lam, const double *__restrict__ sig_c, const double *__restrict__ mz,
               SweepCols sc, double tol, long long max_iter, long long *iters, double *last_err) {
    extern __shared__ __align__(16) double fsm[];
    __shared__ double red[SWF_THREADS / 32];
    const int64_t N = kv.N;
    const int64_t b = blockIdx.x;
    double *cur = fsm, *smat = fsm + ldn;
    {
        double *sm_m = smat;
        for (int m = 0; m < kv.n_modes; ++m) {
            const int n = kv.shape[kv.modes[m].dim];
            fibre2_stage(kv.modes[m], n, sm_m);
            sm_m += fibre2_smat_doubles(kv.modes[m], n);
        }
    }
    const double th = sc.theta[b], g = sc.gamma[b], be = sc.beta[b];
    const double omg = 1.0 - g, inv_th = 1.0 / th;
    double *Wb = W + b * ldw;
    constexpr int U = 4;
    const int64_t step = (int64_t)U * SWF_THREADS;
    for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {          // first contraction input from the stored w
        double v[U], h[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t n = n0 + (int64_t)u * SWF_THREADS;
            v[u] = n < N ? Wb[n] : 1.0;
            h[u] = n < N ? h_lam[n] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t n = n0 + (int64_t)u * SWF_THREADS;
            if (n < N) cur[n] = pow_pos_off(v[u], th, h[u]);
        }
    }
    long long it = 0;
    double e = 0.0;
    for (;;) {
        {
            const double *sm_m = smat;
            __syncthreads();
            for (int m = 0; m < kv.n_modes; ++m) {
                const KronMode &md = kv.modes[m];
                const int n = kv.shape[md.dim];
                fibre2_dispatch(md, n, cur, sm_m);
                sm_m += fibre2_smat_doubles(md, n);
                __syncthreads();
            }
        }
        double emax = 0.0;
        for (int64_t n0 = threadIdx.x; n0 < N; n0 += step) {
            double wo[U], sg[U], zz[U], h[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWF_THREADS;
                wo[u] = n < N ? Wb[n] : 0.0;
                sg[u] = n < N ? sig_c[n] : 0.0;
                zz[u] = n < N ? mz[n] : 0.0;
                h[u] = n < N ? h_lam[n] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t n = n0 + (int64_t)u * SWF_THREADS;
                if (n >= N) continue;
                const double t2 = omg * sg[u];
                const double la = 0.5 * (t2 * t2) + omg * zz[u];
                const double y = 1.0 + be * pow_pos_off(cur[n], inv_th, la);
                Wb[n] = y;
                const double dd = fabs(y - wo[u]);
                emax = (dd != dd || emax != emax) ? dd + emax : fmax(emax, dd);   // NaN propagates
                cur[n] = pow_pos_off(y, th, h[u]);                               // next step's contraction input
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, emax, o);
            emax = (other != other || emax != emax) ? other + emax : fmax(emax, other);
        }
        __syncthreads();                                   // red[] of the previous step consumed
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = emax;
        __syncthreads();
        e = red[0];
        for (int i = 1; i < SWF_THREADS / 32; ++i) e = (red[i] != red[i] || e != e) ? red[i] + e : fmax(e, red[i]);
        e = fabs(e);
        ++it;
        if (!(e > tol) || it >= max_iter) break;
    }
    if (threadIdx.x == 0) { iters[b] = it; last_err[b] = e; }
}

static bool sweep_sa_cols_applicable(const sdfs_op *op, const SweepWork &w) {
    // SDFS_SWEEP_SA_FUSED=0: one launch per step over all columns (the A/B switch)
    static const bool allowed = !(getenv("SDFS_SWEEP_SA_FUSED") && atoi(getenv("SDFS_SWEEP_SA_FUSED")) == 0) &&
                                !(getenv("SDFS_SWEEP_FIBRE2") && atoi(getenv("SDFS_SWEEP_FIBRE2")) == 0);
    return allowed && w.fused && fibre2_all_doubles(op->kv) > 0;
}
static int sweep_sa_cols(sdfs_op *op, SweepWork &w, int64_t B, double *W, double tol, long long max_iter) {
    sdfs_ctx *ctx = op->ctx;
    const size_t smem = sweep_fused_smem(op->kv);
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep_sa_col, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SweepCols sc{w.gamma, w.theta, w.beta, nullptr};
    k_sweep_sa_col<<<(unsigned)B, SWF_THREADS, smem, ctx->stream>>>(op->kv, w.ldw, sweep_fused_ldn(op->kv.N), W, w.hl, w.sc, w.mz, sc,
                                                                   tol, max_iter, w.iters, w.last_err);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

// w <- w - x ; error = max|x| ; outer loop condition (successive_approx rule, solvers.py:34-40)
__global__ void k_swn_outer(int64_t N, int64_t ldw, double tol, long long max_iter, SwnState st, SwnPanels p) {
    __shared__ double sm[SWN_THREADS / 32];
    const int64_t b = blockIdx.x;
    if (!st.out_active[b]) return;
    double m = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += SWN_THREADS) {
        const int64_t e = b * ldw + n;
        const double xn = p.X[e];
        p.W[e] -= xn;
        m = nanmax(m, fabs(xn));
    }
    m = cta_reduce_nanmax(m, sm);
    if (threadIdx.x == 0) {
        st.outer_it[b] += 1;
        st.last_err[b] = m;
        if (!(m > tol) || st.outer_it[b] >= max_iter) {
            st.out_active[b] = 0;
            atomicSub(st.n_out_active, 1);
        }
    }
}

__global__ void k_set_int(int *p, int64_t n, int v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int sdfs_sweep_solve_newton(sdfs_op *op, const double *h_prefs, int64_t B, double w_init, double tol,
                                       int64_t max_iter, double rtol, double atol, int64_t krylov_maxiter,
                                       double *d_W_out, int64_t *h_outer_iters, double *h_final_err,
                                       int64_t *h_inner_total, int64_t *total_gemms) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_sweep_solve_newton: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, h_prefs && d_W_out && B >= 1 && max_iter >= 0);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    SweepWork w;
    int rc = sweep_setup(op, h_prefs, B, false, &w);
    const int64_t N = op->kv.N;
    const long long kmax = krylov_maxiter > 0 ? krylov_maxiter : 10 * N;
    double *pan = nullptr, *sca = nullptr;
    long long *lls = nullptr;
    int *ints = nullptr;
    const size_t pdoubles = (size_t)B * w.ldw;
    SwnPanels p{};
    SwnState st{};
    int64_t gemms = 0;
    if (rc == SDFS_OK) {
        cudaError_t e = cudaMallocAsync(&pan, 12 * pdoubles * sizeof(double), ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(pan, 0, 12 * pdoubles * sizeof(double), ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(&sca, 7 * B * sizeof(double), ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(&lls, 3 * B * sizeof(long long), ctx->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(&ints, (3 * B + 3) * sizeof(int), ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(sca, 0, 7 * B * sizeof(double), ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(lls, 0, 3 * B * sizeof(long long), ctx->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(ints, 0, (3 * B + 3) * sizeof(int), ctx->stream);
        if (e != cudaSuccess) rc = sdfs_set_error(ctx, SDFS_ERR_NOMEM, "sweep newton workspace (%.2f GB): %s", 12 * pdoubles * 8 / 1e9, cudaGetErrorString(e));
    }
    if (rc == SDFS_OK) {
        double **pp[12] = {&p.W, &p.G, &p.C, &p.D, &p.X, &p.R, &p.Rh, &p.Pv, &p.Q, &p.Sv, &p.T, &p.Xin};
        for (int i = 0; i < 12; ++i) *pp[i] = pan + i * pdoubles;
        double **sp[7] = {&st.rho, &st.alpha, &st.omega, &st.rs, &st.rho_next, &st.atol2, &st.last_err};
        for (int i = 0; i < 7; ++i) *sp[i] = sca + i * B;
        st.k = lls; st.outer_it = lls + B; st.inner_total = lls + 2 * B;
        st.exit_early = ints; st.in_active = ints + B; st.out_active = ints + 2 * B;
        st.n_in_active = ints + 3 * B; st.n_out_active = ints + 3 * B + 1;
        const int bgrid = (int)B;
        k_fill_panel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(p.W, N, B, w.ldw, w_init);
        k_set_int<<<(int)((B + 255) / 256), 256, 0, ctx->stream>>>(st.out_active, B, max_iter > 0 ? 1 : 0);
        int nb = max_iter > 0 ? (int)B : 0;
        cudaMemcpyAsync(st.n_out_active, &nb, 4, cudaMemcpyHostToDevice, ctx->stream);
        ctx->launches += 2;
        SweepCols sc{w.gamma, w.theta, w.beta, nullptr};
        // SDFS_SWEEP_INNER_FUSED=0: Krylov phases as separate launches over all columns (round 1; the A/B switch)
        static const bool inner_allowed = !(getenv("SDFS_SWEEP_INNER_FUSED") && atoi(getenv("SDFS_SWEEP_INNER_FUSED")) == 0) &&
                                          !(getenv("SDFS_SWEEP_FIBRE2") && atoi(getenv("SDFS_SWEEP_FIBRE2")) == 0);
        const bool inner_fused = inner_allowed && w.fused && fibre2_all_doubles(op->kv) > 0;
        const size_t fused_smem = sweep_fused_smem(op->kv);
        if (inner_fused)
            cudaFuncSetAttribute(k_swn_inner_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem);
        int n_out = nb;
        auto read_int = [&](const int *d, int *h) -> int {
            cudaError_t e = cudaMemcpyAsync(h, d, 4, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) return sdfs_set_error(ctx, SDFS_ERR_CUDA, "sweep newton: %s", cudaGetErrorString(e));
            return SDFS_OK;
        };
        while (rc == SDFS_OK && n_out > 0) {
            k_swn_prologue<<<bgrid, SWN_THREADS, 0, ctx->stream>>>(N, w.ldw, w.hl, sc, st, p);
            SweepEpi e1{1, p.W, p.G, p.D, nullptr, nullptr, nullptr};
            rc = sweep_PV(op, w, B, p.Xin, e1, sc);
            ++gemms;
            if (rc) break;
            cudaMemsetAsync(st.n_in_active, 0, 4, ctx->stream);
            k_swn_init<<<bgrid, SWN_THREADS, 0, ctx->stream>>>(N, w.ldw, rtol, atol, st, p);
            ctx->launches += 2;
            int n_in = 0;
            if (inner_fused) {
                // every column's whole Krylov solve in one launch (one CTA per column); the application count reported
                // is that of the column that iterated longest
                int *max_k = ints + 3 * B + 2;
                cudaMemsetAsync(max_k, 0, 4, ctx->stream);
                k_swn_inner_fused<<<bgrid, SWF_THREADS, fused_smem, ctx->stream>>>(op->kv, w.ldw, sweep_fused_ldn(N), kmax, st, p, max_k);
                ctx->launches++;
                int mk = 0;
                rc = read_int(max_k, &mk);
                gemms += 2 * (int64_t)mk;
                if (rc == SDFS_OK) {
                    cudaError_t le = cudaGetLastError();
                    if (le != cudaSuccess) rc = sdfs_set_error(ctx, SDFS_ERR_CUDA, "sweep newton (fused Krylov): %s", cudaGetErrorString(le));
                }
            } else {
            rc = read_int(st.n_in_active, &n_in);
            while (rc == SDFS_OK && n_in > 0) {
                k_swn_phase1<<<bgrid, SWN_THREADS, 0, ctx->stream>>>(N, w.ldw, st, p);
                SweepEpi e2{2, nullptr, p.Q, nullptr, p.D, p.Pv, nullptr};
                rc = sweep_PV(op, w, B, p.Xin, e2, sc);
                if (rc) break;
                k_swn_phase3<<<bgrid, SWN_THREADS, 0, ctx->stream>>>(N, w.ldw, st, p);
                SweepEpi e3{2, nullptr, p.T, nullptr, p.D, p.Sv, nullptr};
                rc = sweep_PV(op, w, B, p.Xin, e3, sc);
                if (rc) break;
                k_swn_phase5<<<bgrid, SWN_THREADS, 0, ctx->stream>>>(N, w.ldw, kmax, st, p);
                ctx->launches += 3;
                gemms += 2;
                rc = read_int(st.n_in_active, &n_in);     // 4-byte poll per Krylov iteration (>= 10 ms of GEMM each)
            }
            }
            if (rc) break;
            k_swn_outer<<<bgrid, SWN_THREADS, 0, ctx->stream>>>(N, w.ldw, tol, (long long)max_iter, st, p);
            ctx->launches++;
            rc = read_int(st.n_out_active, &n_out);
        }
    }
    if (rc == SDFS_OK) {
        k_copy_panel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(p.W, w.ldw, d_W_out, N, N, B);
        ctx->launches++;
        std::vector<long long> it(B), inn(B);
        cudaError_t e = cudaMemcpyAsync(it.data(), st.outer_it, B * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(inn.data(), st.inner_total, B * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && h_final_err) e = cudaMemcpyAsync(h_final_err, st.last_err, B * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = sdfs_set_error(ctx, SDFS_ERR_CUDA, "sweep newton: %s", cudaGetErrorString(e));
        for (int64_t b = 0; b < B; ++b) {
            if (h_outer_iters) h_outer_iters[b] = it[b];
            if (h_inner_total) h_inner_total[b] = inn[b];
        }
        if (total_gemms) *total_gemms = gemms;
    }
    cudaStreamSynchronize(ctx->stream);
    if (pan) cudaFreeAsync(pan, ctx->stream);     // stream-ordered pool: the next sweep reuses the blocks
    if (sca) cudaFreeAsync(sca, ctx->stream);
    if (lls) cudaFreeAsync(lls, ctx->stream);
    if (ints) cudaFreeAsync(ints, ctx->stream);
    w.free_all();
    return rc;
}
