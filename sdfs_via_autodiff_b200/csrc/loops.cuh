// Device-resident fixed-point loops: successive approximation and Newton with an
// on-device Krylov solve (BiCGSTAB with JAX's recurrence, or restarted GMRES).
//
// Each solve is ONE cooperative kernel launch: every iteration's operator pass,
// vector updates, reductions and the stopping test run on the device; the host
// blocks once, at the end.  Reductions are deterministic: per-CTA partials land
// in fixed slots and every CTA sums all slots in the same order after the grid
// barrier, so all CTAs (and all ranks) take identical control-flow decisions.
//
// Multi-GPU (one process per GPU, row-sharded P): the same kernels run on every
// rank.  Producers store their slice of the next matvec input and their partial
// sums straight into every peer's exchange arena (NVLink peer stores), and the
// grid barrier is extended by a flag exchange between the ranks -- the
// all-gather is fused into the operator epilogue instead of being a separate
// collective launch.
//
// Reference semantics reproduced: successive_approx (solvers.py:19-48),
// newton_solver (solvers.py:51-95), jax.scipy.sparse.linalg.bicgstab
// (x0 = 0, atol2 = max(tol^2 <b,b>, atol^2), early-exit half step, breakdown codes).
#pragma once
#include "common.cuh"
#include "rowdot.cuh"
#include "cont.cuh"
#include "arena.cuh"

int small_sa_try(sdfs_op *op, const double *d_w_init, double tol, int64_t max_iter, double *d_w_out,
                 double *d_err_hist, int64_t hist_stride, int64_t hist_cap);   // small.cu

#define TRY(x) do { int _rc = (x); if (_rc != SDFS_OK) return _rc; } while (0)

#define HIST_CAP 120       // outer Newton iterations recorded in the status page
#define VU 4               // elements per thread and trip in the Krylov vector phases
#define VUC 8              // ... in the phases that read at most four vectors
#define GMRES_MAX_RESTART 64
#define GMRES_WS_BYTES (((GMRES_MAX_RESTART + 1) * 2 + GMRES_MAX_RESTART * 3 + GMRES_MAX_RESTART * GMRES_MAX_RESTART) * sizeof(double) + 16)

struct LoopStatus {        // lives in ctx->d_status (4 KB)
    long long iters;
    long long inner_total;
    long long matvecs;
    long long abort_code;  // 0 ok, 1 peer barrier timeout
    double final_err;
    unsigned long long epoch_end;   // barrier epoch after the loop (identical on all ranks)
    double pad[2];
    unsigned long long t_apply_ns, n_apply, t_total_ns, t_mode_ns[4], t_epi_ns, t_vec_ns[3], t_red_ns;   // block 0's clock: time inside operator applications (diagnostic)
    double outer_err[HIST_CAP];
    long long inner_iters[HIST_CAP];
};
static_assert(sizeof(LoopStatus) <= 4096 - 512, "status page");

struct LoopEnv {
    int rank, nranks;
    unsigned long long epoch0;
    unsigned long long *flags[SDFS_MAX_RANKS];   // flags[r][src]: arrival counter written by rank src
    double *slots[SDFS_MAX_RANKS];               // slots of rank r: [NSETS][nranks][SDFS_MAX_GRID][NVAL]
    double *xin[SDFS_MAX_RANKS][2];              // matvec input buffers of rank r
    LoopStatus *status;
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- device-side synchronisation and reductions ---------------------------
// Barrier across every CTA of every rank.  Returns false (uniformly) after a peer
// timeout so the kernel can unwind instead of hanging the GPU.
__device__ __forceinline__ bool all_sync(cg::grid_group &grid, const LoopEnv &env, unsigned long long &epoch) {
    // writer side of the generic -> async proxy hand-over: the matvec input stored by this thread
    // is read by the TMA unit (async proxy) of other CTAs after this barrier
    asm volatile("fence.proxy.async;" ::: "memory");
    if (env.nranks == 1) {
        grid.sync();
        return true;
    }
    __threadfence_system();
    grid.sync();
    epoch += 1;
    if (blockIdx.x == 0 && threadIdx.x < env.nranks) {
        const int peer = threadIdx.x;
        st_release_sys(env.flags[peer] + env.rank, epoch);
        const long long t0 = clock64();
        while (ld_acquire_sys(env.flags[env.rank] + peer) < epoch) {
            if (clock64() - t0 > SDFS_PEER_TIMEOUT_CLOCKS) {
                env.status->abort_code = 1;
                break;
            }
        }
    }
    grid.sync();
    return __ldcg(&env.status->abort_code) == 0;
}

// Publish this CTA's partials of reduction set `set` to every rank.
template <int NV>
__device__ __forceinline__ void publish_partials(const LoopEnv &env, int set, const double (&cta_val)[NV]) {
    for (int r = 0; r < env.nranks; ++r) {
        double *dst = env.slots[r] + (((size_t)set * env.nranks + env.rank) * SDFS_MAX_GRID + blockIdx.x) * NVAL;
#pragma unroll
        for (int j = 0; j < NV; ++j) dst[j] = cta_val[j];
    }
}

// CTA-level sum of per-thread partials (fixed order), result valid in thread 0.
template <int NV>
__device__ __forceinline__ void cta_sum(double (&v)[NV], double *smem /* >= SDFS_WARPS*NV */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = warp_sum(v[j]);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < NV; ++j) smem[warp * NV + j] = v[j];
    __syncthreads();
    if (threadIdx.x == 0)
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            double s = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += smem[w * NV + j];
            v[j] = s;
        }
}

// Sum (or NaN-propagating max) of a slot set over all CTAs of all ranks; every
// thread of every CTA receives bit-identical values.
template <int NV, bool IS_MAX>
__device__ __forceinline__ void gather_partials(const LoopEnv &env, int set, double (&out)[NV], double *smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = gridDim.x;
    if (warp == 0) {
        double acc[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) acc[j] = 0.0;
        const int total = env.nranks * G;
        for (int e = lane; e < total; e += 32) {
            const int src = e / G, cta = e % G;
            const double *p = env.slots[env.rank] + (((size_t)set * env.nranks + src) * SDFS_MAX_GRID + cta) * NVAL;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const double x = __ldcg(p + j);
                acc[j] = IS_MAX ? nanmax(acc[j], x) : acc[j] + x;
            }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            acc[j] = IS_MAX ? warp_nanmax(acc[j]) : warp_sum(acc[j]);
            if (lane == 0) smem[j] = acc[j];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; ++j) out[j] = smem[j];
    __syncthreads();
}

template <int NV, bool IS_MAX>
__device__ __forceinline__ bool grid_allreduce(cg::grid_group &grid, const LoopEnv &env, unsigned long long &epoch,
                                               int set, double (&v)[NV], double *smem) {
    if (IS_MAX) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = warp_nanmax(v[j]);
        __syncthreads();
        if (lane == 0)
#pragma unroll
            for (int j = 0; j < NV; ++j) smem[warp * NV + j] = v[j];
        __syncthreads();
        if (threadIdx.x == 0)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                double s = 0.0;
                for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s = nanmax(s, smem[w * NV + j]);
                v[j] = s;
            }
    } else {
        cta_sum<NV>(v, smem);
    }
    if (threadIdx.x == 0) publish_partials<NV>(env, set, v);
    if (!all_sync(grid, env, epoch)) return false;
    gather_partials<NV, IS_MAX>(env, set, v, smem);
    return true;
}

__device__ __forceinline__ void store_all_ranks(const LoopEnv &env, int buf, int64_t n, double val) {
    if (env.nranks == 1) { env.xin[0][buf][n] = val; return; }      // single GPU: no loop, no pointer table walk
#pragma unroll 1
    for (int r = 0; r < env.nranks; ++r) env.xin[r][buf][n] = val;
}

// ---- operator abstraction ---------------------------------------------------
// apply(view, xin, epi): epi(n, s) for every row n owned by this rank, s = (P xin)[n].
struct Scratch {
    RowPipe<1> *rp;        // TMA ring in dynamic shared memory (dense operators)
    PipeState st;
    double *smat;          // factor matrix staging (factor-form operators), dynamic shared memory
    double *stage;         // per-warp fragment stage of the tensor-core contraction (rowdot.cuh), after smat
};

struct DenseLoopOp {
    static constexpr int kMinBlocks = 1;
    static constexpr int kThreads = SDFS_THREADS;      // 8 consumer warps + the TMA producer warp
    static constexpr bool kTwoStage = false;
    DenseView dv;
    __host__ __device__ size_t dyn_smem() const { return dv.vec2 ? sizeof(RowPipe<1>) : 0; }
    __device__ __forceinline__ void init(Scratch &sc) const {
        if (dv.vec2) pipe_init(sc.rp, sc.st);
    }
    __device__ __forceinline__ int64_t N() const { return dv.N; }
    __device__ __forceinline__ int64_t row_begin() const { return dv.row_begin; }
    __device__ __forceinline__ int64_t row_end() const { return dv.row_end; }
    __device__ __forceinline__ const double *a_row() const { return dv.a_row; }
    __device__ __forceinline__ const double *a_col() const { return dv.a_col; }
    __device__ __forceinline__ double beta() const { return dv.beta; }
    __device__ __forceinline__ double theta() const { return dv.theta; }
    // hooks of the loop kernels: what is staged as the operator input, the column scaling of the
    // linearised map, the row factor of d = beta s^((1-theta)/theta) rowfac
    static constexpr bool kNeedsW = false;
    __device__ __forceinline__ double stage_T(int64_t n, double w) const { return dv.a_col[n] * pow(w, dv.theta); }
    __device__ __forceinline__ double stage_c(int64_t n, double w) const { return dv.a_col[n] * pow(w, dv.theta - 1.0); }
    __device__ __forceinline__ double rowfac(int64_t n) const { return dv.a_row[n]; }
    // linear map on the staged vector: epi(n, (P xin)[n])
    template <class Epi>
    __device__ __forceinline__ bool apply(cg::grid_group &, const LoopEnv &, unsigned long long &, Scratch &sc,
                                          const double *xin, const double *, Epi &&epi) const {
        dense_pass<1>(dv, xin, xin, sc.rp, sc.st, [&](int64_t n, double s0, double) { epi(n, s0); });
        return true;
    }
    // T pass: epi(n, s) with T w = 1 + beta s^(1/theta)
    template <class Epi>
    __device__ __forceinline__ bool apply_T(cg::grid_group &, const LoopEnv &, unsigned long long &, Scratch &sc,
                                            const double *xin, Epi &&epi) const {
        dense_pass<1>(dv, xin, xin, sc.rp, sc.st, [&](int64_t n, double s0, double) { epi(n, dv.a_row[n] * s0); });
        return true;
    }
};

struct KronLoopOp {
    static constexpr int kMinBlocks = 2;     // two CTAs per SM
    static constexpr int kThreads = 256;     // 128 registers per thread: at 288 threads (<= 112, 96 in practice) the Krylov vector
                                             // phases spilled their batched loads to local memory (STL/LDL in the SASS)
    KronView kv;
    double *tmp0, *tmp1;
    __host__ __device__ size_t dyn_smem() const { return (KRON_SMAT_DOUBLES + (kThreads / 32) * KRON_STAGE_DOUBLES_PER_WARP) * sizeof(double); }
    __device__ __forceinline__ void init(Scratch &) const {}
    __device__ __forceinline__ int64_t N() const { return kv.N; }
    __device__ __forceinline__ int64_t row_begin() const { return kv.row_begin; }     // slab of the leading axis (sharded view)
    __device__ __forceinline__ int64_t row_end() const { return kv.row_end; }
    __device__ __forceinline__ const double *a_row() const { return kv.a_row; }
    __device__ __forceinline__ const double *a_col() const { return kv.a_col; }
    __device__ __forceinline__ double beta() const { return kv.beta; }
    __device__ __forceinline__ double theta() const { return kv.theta; }
    static constexpr bool kNeedsW = false;
    __device__ __forceinline__ double stage_T(int64_t n, double w) const { return kv.a_col[n] * pow_pos(w, kv.theta); }
    __device__ __forceinline__ double stage_c(int64_t n, double w) const { return kv.a_col[n] * pow_pos(w, kv.theta - 1.0); }
    __device__ __forceinline__ double rowfac(int64_t n) const { return kv.a_row[n]; }
    template <class Epi>
    __device__ __forceinline__ bool apply_T(cg::grid_group &grid, const LoopEnv &env, unsigned long long &epoch,
                                            Scratch &sc, const double *xin, Epi &&epi) const {
        return apply(grid, env, epoch, sc, xin, nullptr, [&](int64_t n, double s) { epi(n, kv.a_row[n] * s); });
    }
    // Every mode is contracted by the one out-of-line storing function (exact tile counts, cp.async-staged
    // fragments), the barriers between modes are LOCAL grid barriers (only the first mode reads data written
    // by other ranks, and that was ordered by the all_sync before the call), and the epilogue runs as a
    // coalesced vector phase over the finished contraction.  (Fused into the last contraction's sink, as in
    // round 1, the epilogue's scattered dependent loads - d, p, r-hat per output - made that mode cost 169 us
    // against 68 us for the others at 9.8 M states, and every call site carried its own copy of the
    // tensor-core code.)
    static constexpr bool kTwoStage = true;      // contract() leaves P xin in a vector; callers may run their own vector phase on it
    __device__ __forceinline__ const double *contract(cg::grid_group &grid, const LoopEnv &env, Scratch &sc, const double *xin) const {
        const double *in = xin;
        const bool clk = blockIdx.x == 0 && threadIdx.x == 0;
        unsigned long long t0 = 0, t1;
        if (clk) t0 = gtimer();
        const unsigned long long tb = t0;
        for (int m = 0; m < kv.n_modes; ++m) {
            double *out = (m & 1) ? tmp1 : tmp0;
            kron_mode_store(kv, m, in, out, sc.smat, sc.stage);
            grid.sync();                        // (orders the stores: cooperative-groups grid barriers fence)
            if (clk) { t1 = gtimer(); if (m < 4) env.status->t_mode_ns[m] += t1 - t0; t0 = t1; }
            in = out;
        }
        if (clk) { env.status->t_apply_ns += t0 - tb; env.status->n_apply += 1; }
        return in;
    }
    template <class Epi>
    __device__ __forceinline__ bool apply(cg::grid_group &grid, const LoopEnv &env, unsigned long long &epoch,
                                          Scratch &sc, const double *xin, const double *, Epi &&epi) const {
        const double *in = contract(grid, env, sc, xin);
        const bool clk = blockIdx.x == 0 && threadIdx.x == 0;
        unsigned long long t0 = 0;
        if (clk) t0 = gtimer();
        const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t nth = (int64_t)gridDim.x * blockDim.x;
        const int64_t rb = kv.row_begin, re = kv.row_end;
        for (int64_t n0 = rb + tid; n0 < re; n0 += VU * nth) {
            double sv[VU];
#pragma unroll
            for (int u = 0; u < VU; ++u) {
                const int64_t n = n0 + u * nth;
                sv[u] = n < re ? in[n] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < VU; ++u) {
                const int64_t n = n0 + u * nth;
                if (n < re) epi(n, sv[u]);
            }
        }
        if (clk) env.status->t_epi_ns += gtimer() - t0;
        (void)epoch;
        return true;
    }
};

// Continuous-state operator (cont.cuh): the staged vector is w itself, the linearised map needs
// the full current w beside the vector it is applied to.
struct ContLoopOp {
    static constexpr int kMinBlocks = 2;
    static constexpr int kThreads = SDFS_THREADS;
    static constexpr bool kTwoStage = false;
    static constexpr bool kNeedsW = true;
    ContView cv;
    __host__ __device__ size_t dyn_smem() const { return 0; }
    __device__ __forceinline__ void init(Scratch &) const {}
    __device__ __forceinline__ int64_t N() const { return cv.N; }
    __device__ __forceinline__ int64_t row_begin() const { return cv.row_begin; }
    __device__ __forceinline__ int64_t row_end() const { return cv.row_end; }
    __device__ __forceinline__ double beta() const { return cv.beta; }
    __device__ __forceinline__ double theta() const { return cv.theta; }
    __device__ __forceinline__ double stage_T(int64_t, double w) const { return w; }
    __device__ __forceinline__ double stage_c(int64_t, double) const { return 1.0; }
    __device__ __forceinline__ double rowfac(int64_t n) const { return cont_rowfac(cv, n); }
    template <class Epi>
    __device__ __forceinline__ bool apply_T(cg::grid_group &, const LoopEnv &, unsigned long long &, Scratch &,
                                            const double *xin, Epi &&epi) const {
        cont_pass<0>(cv, xin, nullptr, [&](int64_t n, double kg, double) { epi(n, kg); });
        return true;
    }
    template <class Epi>
    __device__ __forceinline__ bool apply(cg::grid_group &, const LoopEnv &, unsigned long long &, Scratch &,
                                          const double *xin, const double *wfull, Epi &&epi) const {
        cont_pass<1>(cv, wfull, xin, [&](int64_t n, double, double l) { epi(n, l); });
        return true;
    }
};

// ---------------------------------------------------------------------------
// Successive approximation
// ---------------------------------------------------------------------------
struct SAArgs {
    const double *w_init;
    double *w[2];          // ping-pong iterates (full length; each rank fills its rows)
    double *w_out;
    double tol;
    long long max_iter;
    double *err_hist;
    long long hist_stride, hist_cap;
};

template <class Op>
__global__ void __launch_bounds__(Op::kThreads, Op::kMinBlocks) k_sa_loop(const __grid_constant__ Op op, const __grid_constant__ SAArgs a, const __grid_constant__ LoopEnv env) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ double smem[SDFS_WARPS * NVAL + NVAL];
    Scratch sc;
    sc.rp = reinterpret_cast<RowPipe<1> *>(dyn_smem);
    sc.smat = reinterpret_cast<double *>(dyn_smem);
    sc.stage = sc.smat + KRON_SMAT_DOUBLES;
    op.init(sc);
    unsigned long long epoch = env.epoch0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t rb = op.row_begin(), re = op.row_end();
    const double beta = op.beta(), inv_theta = 1.0 / op.theta();

    // stage the operator input of w0 for own rows, broadcast to every rank; w[0] = w0
    for (int64_t n = rb + tid; n < re; n += nth) {
        const double w0 = a.w_init[n];
        a.w[0][n] = w0;
        store_all_ranks(env, 0, n, op.stage_T(n, w0));
    }
    if (!all_sync(grid, env, epoch)) return;

    long long it = 0;
    double error = a.tol + 1.0;
    while (error > a.tol && it < a.max_iter) {
        const int cur = (int)(it & 1), nxt = cur ^ 1;
        const double *w_cur = a.w[cur];
        double *w_nxt = a.w[nxt];
        double part[1] = {0.0};
        bool ok = op.apply_T(grid, env, epoch, sc, env.xin[env.rank][cur], [&](int64_t n, double s) {
            const double y = 1.0 + beta * pow(s, inv_theta);
            w_nxt[n] = y;
            store_all_ranks(env, nxt, n, op.stage_T(n, y));
            part[0] = nanmax(part[0], fabs(y - w_cur[n]));
        });
        if (!ok) return;
        if (!grid_allreduce<1, true>(grid, env, epoch, (int)(it & 1), part, smem)) return;
        error = part[0];
        if (tid == 0 && a.err_hist && (it % a.hist_stride) == 0 && (it / a.hist_stride) < a.hist_cap)   // every rank: its own buffer, identical values
            a.err_hist[it / a.hist_stride] = error;
        ++it;
    }
    const double *w_fin = a.w[it & 1];
    for (int64_t n = rb + tid; n < re; n += nth) a.w_out[n] = w_fin[n];
    if (tid == 0) {
        env.status->iters = it;
        env.status->final_err = error;
        env.status->epoch_end = epoch;
    }
}

// ---------------------------------------------------------------------------
// Newton with on-device Krylov solve
// ---------------------------------------------------------------------------
struct NewtonArgs {
    const double *w_init;
    double *w_out;
    // N-vectors (full length allocations; each rank touches its own rows)
    double *__restrict__ w, *__restrict__ g, *__restrict__ c, *__restrict__ d, *__restrict__ x, *__restrict__ r,
        *__restrict__ rhat, *__restrict__ p, *__restrict__ q, *__restrict__ s, *__restrict__ t;   // distinct work vectors
    double *V;             // GMRES basis: (restart+1) vectors of ldv doubles
    long long ldv;
    double tol;
    long long max_iter;
    int krylov;
    double rtol, atol;
    int restart;
    long long krylov_maxiter;
};

// reduction slot sets (each phase owns one, so a set is rewritten only after at
// least one other barrier has passed)
enum { SET_B = 0, SET_D = 1, SET_E = 2, SET_F = 3, SET_G = 4, SET_H = 5, SET_X = 6, SET_Y = 7 };

template <class Op>
__device__ __forceinline__ bool bicgstab_device(cg::grid_group &grid, const Op &op, const NewtonArgs &a,
                                                const LoopEnv &env, unsigned long long &epoch, Scratch &sc, double *smem,
                                                double bs, long long &k_out, long long &matvecs) {
    // on entry: r = rhat = p = q = g (= b), x = 0; bs = <b,b>
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t rb = op.row_begin(), re = op.row_end();
    const double atol2 = fmax(a.rtol * a.rtol * bs, a.atol * a.atol);
    double rs = bs;          // <r,r>
    double rho_next = bs;    // <rhat,r>
    double alpha = 1.0, omega = 1.0, rho = 1.0;
    long long k = 0;
    while (rs > atol2 && k < a.krylov_maxiter && k >= 0) {
        // C: p = r + beta (p - omega q); xin = c .* p
        const double rho_ = rho_next;
        const double beta = rho_ / rho * alpha / omega;
        const bool clk = blockIdx.x == 0 && threadIdx.x == 0;
        unsigned long long tp = 0, tq = 0;
        if (clk) tp = gtimer();
        // (vector phases: VU independent elements per trip, every load issued before the first store - one
        // element per trip left the loops latency-bound at ~half of HBM bandwidth with 18 warps per SM)
        for (int64_t n0 = rb + tid; n0 < re; n0 += VUC * nth) {
            double rv[VUC], pv[VUC], qv[VUC], cv[VUC];
#pragma unroll
            for (int u = 0; u < VUC; ++u) {
                const int64_t n = n0 + u * nth;
                if (n < re) { rv[u] = a.r[n]; pv[u] = a.p[n]; qv[u] = a.q[n]; cv[u] = a.c[n]; }
            }
#pragma unroll
            for (int u = 0; u < VUC; ++u) {
                const int64_t n = n0 + u * nth;
                if (n < re) {
                    const double pn = rv[u] + beta * (pv[u] - omega * qv[u]);
                    a.p[n] = pn;
                    store_all_ranks(env, 0, n, cv[u] * pn);
                }
            }
        }
        if (!all_sync(grid, env, epoch)) return false;
        if (clk) { tq = gtimer(); env.status->t_vec_ns[0] += tq - tp; }
        // D: q = J p = d .* P(c .* p) - p ; <rhat,q>
        double v1[1] = {0.0};
        if constexpr (Op::kTwoStage) {
            const double *sum = op.contract(grid, env, sc, env.xin[env.rank][0]);
            unsigned long long te = 0;
            if (clk) te = gtimer();
            for (int64_t n0 = rb + tid; n0 < re; n0 += VU * nth) {
                double sv[VU], dv[VU], pv[VU], hv[VU];
#pragma unroll
                for (int u = 0; u < VU; ++u) {
                    const int64_t n = n0 + u * nth;
                    if (n < re) { sv[u] = sum[n]; dv[u] = a.d[n]; pv[u] = a.p[n]; hv[u] = a.rhat[n]; }
                }
#pragma unroll
                for (int u = 0; u < VU; ++u) {
                    const int64_t n = n0 + u * nth;
                    if (n < re) {
                        const double qn = dv[u] * sv[u] - pv[u];
                        a.q[n] = qn;
                        v1[0] += hv[u] * qn;
                    }
                }
            }
            if (clk) env.status->t_epi_ns += gtimer() - te;
        } else if (!op.apply(grid, env, epoch, sc, env.xin[env.rank][0], env.xin[env.rank][1], [&](int64_t n, double sum) {
                const double qn = a.d[n] * sum - a.p[n];
                a.q[n] = qn;
                v1[0] += a.rhat[n] * qn;
            })) return false;
        if (clk) tp = gtimer();
        if (!grid_allreduce<1, false>(grid, env, epoch, SET_D, v1, smem)) return false;
        if (clk) { tq = gtimer(); env.status->t_red_ns += tq - tp; tp = tq; }
        const double alpha_ = rho_ / v1[0];
        // E: s = r - alpha q ; <s,s> ; xin = c .* s
        double v2[1] = {0.0};
        for (int64_t n0 = rb + tid; n0 < re; n0 += VUC * nth) {
            double rv[VUC], qv[VUC], cv[VUC];
#pragma unroll
            for (int u = 0; u < VUC; ++u) {
                const int64_t n = n0 + u * nth;
                if (n < re) { rv[u] = a.r[n]; qv[u] = a.q[n]; cv[u] = a.c[n]; }
            }
#pragma unroll
            for (int u = 0; u < VUC; ++u) {
                const int64_t n = n0 + u * nth;
                if (n < re) {
                    const double sn = rv[u] - alpha_ * qv[u];
                    a.s[n] = sn;
                    v2[0] += sn * sn;
                    store_all_ranks(env, 0, n, cv[u] * sn);
                }
            }
        }
        if (!grid_allreduce<1, false>(grid, env, epoch, SET_E, v2, smem)) return false;
        if (clk) { tq = gtimer(); env.status->t_vec_ns[1] += tq - tp; }
        const bool exit_early = v2[0] < atol2;
        // F: t = J s ; <t,s>, <t,t>
        double v3[2] = {0.0, 0.0};
        if constexpr (Op::kTwoStage) {
            const double *sum = op.contract(grid, env, sc, env.xin[env.rank][0]);
            unsigned long long te = 0;
            if (clk) te = gtimer();
            for (int64_t n0 = rb + tid; n0 < re; n0 += VU * nth) {
                double sv[VU], dv[VU], xv[VU];
#pragma unroll
                for (int u = 0; u < VU; ++u) {
                    const int64_t n = n0 + u * nth;
                    if (n < re) { sv[u] = sum[n]; dv[u] = a.d[n]; xv[u] = a.s[n]; }
                }
#pragma unroll
                for (int u = 0; u < VU; ++u) {
                    const int64_t n = n0 + u * nth;
                    if (n < re) {
                        const double tn = dv[u] * sv[u] - xv[u];
                        a.t[n] = tn;
                        v3[0] += tn * xv[u];
                        v3[1] += tn * tn;
                    }
                }
            }
            if (clk) env.status->t_epi_ns += gtimer() - te;
        } else if (!op.apply(grid, env, epoch, sc, env.xin[env.rank][0], env.xin[env.rank][1], [&](int64_t n, double sum) {
                const double sn = a.s[n];
                const double tn = a.d[n] * sum - sn;
                a.t[n] = tn;
                v3[0] += tn * sn;
                v3[1] += tn * tn;
            })) return false;
        if (!grid_allreduce<2, false>(grid, env, epoch, SET_F, v3, smem)) return false;
        if (clk) tp = gtimer();
        const double omega_ = v3[0] / v3[1];
        matvecs += 2;
        // G: x, r updates ; <r,r>, <rhat,r>
        double v4[2] = {0.0, 0.0};
        for (int64_t n0 = rb + tid; n0 < re; n0 += VU * nth) {
            double pv[VU], sv[VU], xv[VU], tv[VU], hv[VU];
#pragma unroll
            for (int u = 0; u < VU; ++u) {
                const int64_t n = n0 + u * nth;
                if (n < re) { pv[u] = a.p[n]; sv[u] = a.s[n]; xv[u] = a.x[n]; tv[u] = a.t[n]; hv[u] = a.rhat[n]; }
            }
#pragma unroll
            for (int u = 0; u < VU; ++u) {
                const int64_t n = n0 + u * nth;
                if (n < re) {
                    const double pn = pv[u], sn = sv[u];
                    double xn, rn;
                    if (exit_early) {
                        xn = xv[u] + alpha_ * pn;
                        rn = sn;
                    } else {
                        xn = xv[u] + (alpha_ * pn + omega_ * sn);
                        rn = sn - omega_ * tv[u];
                    }
                    a.x[n] = xn;
                    a.r[n] = rn;
                    v4[0] += rn * rn;
                    v4[1] += hv[u] * rn;
                }
            }
        }
        if (!grid_allreduce<2, false>(grid, env, epoch, SET_G, v4, smem)) return false;
        if (clk) { tq = gtimer(); env.status->t_vec_ns[2] += tq - tp; }
        rs = v4[0];
        rho_next = v4[1];
        long long k_ = (omega_ == 0.0 || alpha_ == 0.0) ? -11 : k + 1;
        if (rho_ == 0.0) k_ = -10;
        k = k_;
        alpha = alpha_;
        omega = omega_;
        rho = rho_;
    }
    k_out = k;
    return true;
}

// Restarted GMRES(m), classical Gram-Schmidt applied twice, Givens rotations;
// x0 = 0; stops when |residual estimate| <= max(rtol ||b||, atol).
template <class Op>
__device__ __forceinline__ bool gmres_device(cg::grid_group &grid, const Op &op, const NewtonArgs &a,
                                             const LoopEnv &env, unsigned long long &epoch, Scratch &sc, double *smem,
                                             double *hs /* shared: Hessenberg workspace */, double bs,
                                             long long &k_out, long long &matvecs) {
    // on entry: r = g (= b), x = 0
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t rb = op.row_begin(), re = op.row_end();
    const int m = a.restart;
    const double target = fmax(a.rtol * sqrt(bs), a.atol);
    double beta = sqrt(bs);
    long long its = 0;
    unsigned gs_count = 0;
    // shared small arrays (identical in every CTA): H column h[0..m], cs, sn, gvec, and R (upper triangular)
    double *Hc = hs;                       // m+1
    double *cs = Hc + (GMRES_MAX_RESTART + 1);
    double *sn = cs + GMRES_MAX_RESTART;
    double *gv = sn + GMRES_MAX_RESTART;   // m+1
    double *Rm = gv + (GMRES_MAX_RESTART + 1);   // m x m upper triangular, row-major stride m
    double *yv = Rm + GMRES_MAX_RESTART * GMRES_MAX_RESTART;   // m
    while (beta > target && its < a.krylov_maxiter) {
        // V0 = r / beta ; xin = c .* V0
        for (int64_t n = rb + tid; n < re; n += nth) {
            const double v = a.r[n] / beta;
            a.V[n] = v;
            store_all_ranks(env, 0, n, a.c[n] * v);
        }
        if (threadIdx.x == 0) gv[0] = beta;
        if (!all_sync(grid, env, epoch)) return false;
        int j_used = 0;
        double res = beta;
        for (int j = 0; j < m; ++j) {
            double *Vj = a.V + (long long)j * a.ldv;
            double *Wv = a.V + (long long)(j + 1) * a.ldv;
            // w = J V_j
            if (!op.apply(grid, env, epoch, sc, env.xin[env.rank][0], env.xin[env.rank][1], [&](int64_t n, double sum) {
                    Wv[n] = a.d[n] * sum - Vj[n];
                })) return false;
            its += 1;
            matvecs += 1;
            if (!all_sync(grid, env, epoch)) return false;
            // classical Gram-Schmidt applied twice, 8 basis vectors per barrier:
            // h = V^T w ; w -= V h
            for (int pass = 0; pass < 2; ++pass) {
                for (int i0 = 0; i0 <= j; i0 += NVAL) {
                    const int nb = (j + 1 - i0) < NVAL ? (j + 1 - i0) : NVAL;
                    double hv[NVAL];
#pragma unroll
                    for (int b = 0; b < NVAL; ++b) hv[b] = 0.0;
                    const double *Vb = a.V + (long long)i0 * a.ldv;
                    for (int64_t n = rb + tid; n < re; n += nth) {
                        const double wn = Wv[n];
#pragma unroll
                        for (int b = 0; b < NVAL; ++b)
                            if (b < nb) hv[b] += Vb[(long long)b * a.ldv + n] * wn;
                    }
                    if (!grid_allreduce<NVAL, false>(grid, env, epoch, (gs_count++ & 1) ? SET_X : SET_Y, hv, smem)) return false;
                    for (int64_t n = rb + tid; n < re; n += nth) {
                        double wn = Wv[n];
#pragma unroll
                        for (int b = 0; b < NVAL; ++b)
                            if (b < nb) wn -= hv[b] * Vb[(long long)b * a.ldv + n];
                        Wv[n] = wn;
                    }
                    if (threadIdx.x == 0) {
#pragma unroll
                        for (int b = 0; b < NVAL; ++b)
                            if (b < nb) {
                                if (pass == 0) Hc[i0 + b] = hv[b];
                                else Hc[i0 + b] += hv[b];
                            }
                    }
                    // the update of Wv is elementwise on own rows: no barrier needed before the next dots
                }
            }
            double nv[1] = {0.0};
            for (int64_t n = rb + tid; n < re; n += nth) nv[0] += Wv[n] * Wv[n];
            if (!grid_allreduce<1, false>(grid, env, epoch, SET_E, nv, smem)) return false;
            const double hnext = sqrt(nv[0]);
            // normalise and stage next matvec input
            if (hnext != 0.0 && j + 1 < m) {
                for (int64_t n = rb + tid; n < re; n += nth) {
                    const double v = Wv[n] / hnext;
                    Wv[n] = v;
                    store_all_ranks(env, 0, n, a.c[n] * v);
                }
            }
            // small dense work, replicated in every CTA by thread 0
            __syncthreads();
            if (threadIdx.x == 0) {
                Hc[j + 1] = hnext;
                for (int i = 0; i < j; ++i) {
                    const double tmp = cs[i] * Hc[i] + sn[i] * Hc[i + 1];
                    Hc[i + 1] = -sn[i] * Hc[i] + cs[i] * Hc[i + 1];
                    Hc[i] = tmp;
                }
                const double den = hypot(Hc[j], Hc[j + 1]);
                cs[j] = Hc[j] / den;
                sn[j] = Hc[j + 1] / den;
                Hc[j] = den;
                gv[j + 1] = -sn[j] * gv[j];
                gv[j] = cs[j] * gv[j];
                for (int i = 0; i <= j; ++i) Rm[i * m + j] = Hc[i];
                smem[0] = fabs(gv[j + 1]);
            }
            __syncthreads();
            res = smem[0];
            __syncthreads();
            j_used = j + 1;
            if (!all_sync(grid, env, epoch)) return false;
            if (res <= target || its >= a.krylov_maxiter) break;
        }
        // back substitution (thread 0 of every CTA), then x += V y
        if (threadIdx.x == 0) {
            for (int i = j_used - 1; i >= 0; --i) {
                double acc = gv[i];
                for (int l = i + 1; l < j_used; ++l) acc -= Rm[i * m + l] * yv[l];
                yv[i] = acc / Rm[i * m + i];
            }
        }
        __syncthreads();
        for (int64_t n = rb + tid; n < re; n += nth) {
            double acc = a.x[n];
            for (int i = 0; i < j_used; ++i) acc += a.V[(long long)i * a.ldv + n] * yv[i];
            a.x[n] = acc;
            store_all_ranks(env, 0, n, a.c[n] * acc);
        }
        if (!all_sync(grid, env, epoch)) return false;
        // true residual r = b - J x
        double rv[1] = {0.0};
        if (!op.apply(grid, env, epoch, sc, env.xin[env.rank][0], env.xin[env.rank][1], [&](int64_t n, double sum) {
                const double rn = a.g[n] - (a.d[n] * sum - a.x[n]);
                a.r[n] = rn;
                rv[0] += rn * rn;
            })) return false;
        matvecs += 1;
        if (!grid_allreduce<1, false>(grid, env, epoch, SET_G, rv, smem)) return false;
        beta = sqrt(rv[0]);
    }
    k_out = its;
    return true;
}

template <class Op>
__global__ void __launch_bounds__(Op::kThreads, Op::kMinBlocks) k_newton_loop(const __grid_constant__ Op op, const __grid_constant__ NewtonArgs a, const __grid_constant__ LoopEnv env) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ double smem[SDFS_WARPS * NVAL + NVAL];
    Scratch sc;
    sc.rp = reinterpret_cast<RowPipe<1> *>(dyn_smem);
    sc.smat = reinterpret_cast<double *>(dyn_smem);
    sc.stage = sc.smat + KRON_SMAT_DOUBLES;
    op.init(sc);
    // GMRES Hessenberg workspace: dynamic shared memory behind the operator's region, present only when the
    // host selected GMRES (35 KB that would otherwise cost the factor-form loops their second CTA per SM)
    double *hs = reinterpret_cast<double *>(dyn_smem + ((op.dyn_smem() + 15) & ~(size_t)15));
    unsigned long long epoch = env.epoch0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t rb = op.row_begin(), re = op.row_end();
    const double theta = op.theta(), beta = op.beta();
    const double inv_theta = 1.0 / theta, d_exp = (1.0 - theta) / theta;

    // A: w = w0 ; staged operator input ; column scaling c of the linearised map
    for (int64_t n = rb + tid; n < re; n += nth) {
        const double w0 = a.w_init[n];
        a.w[n] = w0;
        a.c[n] = op.stage_c(n, w0);
        store_all_ranks(env, 0, n, op.stage_T(n, w0));
        if (Op::kNeedsW) store_all_ranks(env, 1, n, w0);
    }
    if (!all_sync(grid, env, epoch)) return;

    long long it = 0, inner_total = 0, matvecs = 0;
    double error = a.tol + 1.0;
    const unsigned long long t_begin = gtimer();
    while (error > a.tol && it < a.max_iter) {
        // B: s = a_row P xin ; Tw ; g = Tw - w ; d ; Krylov init ; <g,g>
        double vb[1] = {0.0};
        if (!op.apply_T(grid, env, epoch, sc, env.xin[env.rank][0], [&](int64_t n, double sv) {
                const double gn = (1.0 + beta * pow(sv, inv_theta)) - a.w[n];
                a.g[n] = gn;
                a.d[n] = beta * pow(sv, d_exp) * op.rowfac(n);
                a.r[n] = gn; a.rhat[n] = gn; a.p[n] = gn; a.q[n] = gn;
                a.x[n] = 0.0;
                vb[0] += gn * gn;
            })) return;
        matvecs += 1;
        if (!grid_allreduce<1, false>(grid, env, epoch, SET_B, vb, smem)) return;
        long long k_inner = 0;
        if (a.krylov == SDFS_KRYLOV_BICGSTAB) {
            if (!bicgstab_device(grid, op, a, env, epoch, sc, smem, vb[0], k_inner, matvecs)) return;
        } else {
            if (!gmres_device(grid, op, a, env, epoch, sc, smem, hs, vb[0], k_inner, matvecs)) return;
        }
        inner_total += (k_inner > 0 ? k_inner : 0);
        // H: w <- w - x ; error = max|x| ; next xin, c
        double vh[1] = {0.0};
        for (int64_t n = rb + tid; n < re; n += nth) {
            const double xn = a.x[n];
            const double wn = a.w[n] - xn;
            a.w[n] = wn;
            vh[0] = nanmax(vh[0], fabs(xn));
            a.c[n] = op.stage_c(n, wn);
            store_all_ranks(env, 0, n, op.stage_T(n, wn));
            if (Op::kNeedsW) store_all_ranks(env, 1, n, wn);
        }
        if (!grid_allreduce<1, true>(grid, env, epoch, SET_H, vh, smem)) return;
        error = vh[0];
        if (tid == 0 && it < HIST_CAP) {
            env.status->outer_err[it] = error;
            env.status->inner_iters[it] = k_inner;
        }
        ++it;
    }
    for (int64_t n = rb + tid; n < re; n += nth) a.w_out[n] = a.w[n];
    if (tid == 0) {
        env.status->iters = it;
        env.status->final_err = error;
        env.status->inner_total = inner_total;
        env.status->matvecs = matvecs;
        env.status->t_total_ns = gtimer() - t_begin;
        env.status->epoch_end = epoch;
    }
}

// ---------------------------------------------------------------------------
// Anderson acceleration (solvers.py:98-124: jaxopt.AndersonAcceleration, history 10, mixing
// frequency 4, beta 8, ridge 1e-6).  jaxopt is un-vendored: the update rule is restated
// (oracle/solvers.py::anderson_solver) -- parity unpinned.
// ---------------------------------------------------------------------------
#define AND_MAX_HIST 16
struct AndersonArgs {
    const double *w_init;
    double *w_out;
    double *x, *fx;        // current iterate and f(x) (own rows)
    double *X, *R;         // histories: m vectors of ldv doubles each
    long long ldv;
    double tol;
    long long max_iter;
    int m, mix;
    double beta_mix, ridge;
};

// Solve the (m+1)x(m+1) system [[0,1^T],[1,G+ridge I]] [nu;alpha] = e0 (partial pivoting).
__device__ __forceinline__ void anderson_alphas(const double *G, int m, double ridge, double *alpha /* m */) {
    double H[(AND_MAX_HIST + 1) * (AND_MAX_HIST + 2)];
    const int n = m + 1, ldh = n + 1;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= n; ++j) H[i * ldh + j] = 0.0;
    for (int j = 1; j < n; ++j) { H[j] = 1.0; H[j * ldh] = 1.0; }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) H[(i + 1) * ldh + j + 1] = G[i * AND_MAX_HIST + j] + (i == j ? ridge : 0.0);
    H[n] = 1.0;    // right-hand side e0
    for (int c = 0; c < n; ++c) {
        int piv = c;
        double best = fabs(H[c * ldh + c]);
        for (int r = c + 1; r < n; ++r)
            if (fabs(H[r * ldh + c]) > best) { best = fabs(H[r * ldh + c]); piv = r; }
        if (piv != c)
            for (int j = 0; j <= n; ++j) { const double t = H[c * ldh + j]; H[c * ldh + j] = H[piv * ldh + j]; H[piv * ldh + j] = t; }
        const double d = H[c * ldh + c];
        for (int r = c + 1; r < n; ++r) {
            const double fct = H[r * ldh + c] / d;
            for (int j = c; j <= n; ++j) H[r * ldh + j] -= fct * H[c * ldh + j];
        }
    }
    double sol[AND_MAX_HIST + 1];
    for (int r = n - 1; r >= 0; --r) {
        double acc = H[r * ldh + n];
        for (int j = r + 1; j < n; ++j) acc -= H[r * ldh + j] * sol[j];
        sol[r] = acc / H[r * ldh + r];
    }
    for (int i = 0; i < m; ++i) alpha[i] = sol[i + 1];
}

template <class Op>
__global__ void __launch_bounds__(Op::kThreads, Op::kMinBlocks) k_anderson_loop(const __grid_constant__ Op op, const __grid_constant__ AndersonArgs a, const __grid_constant__ LoopEnv env) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ double smem[SDFS_WARPS * NVAL + NVAL];
    __shared__ double sG[AND_MAX_HIST * AND_MAX_HIST];
    __shared__ double sAlpha[AND_MAX_HIST];
    Scratch sc;
    sc.rp = reinterpret_cast<RowPipe<1> *>(dyn_smem);
    sc.smat = reinterpret_cast<double *>(dyn_smem);
    sc.stage = sc.smat + KRON_SMAT_DOUBLES;
    op.init(sc);
    unsigned long long epoch = env.epoch0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t rb = op.row_begin(), re = op.row_end();
    const double beta = op.beta(), inv_theta = 1.0 / op.theta();
    const int m = a.m;
    for (int e = threadIdx.x; e < AND_MAX_HIST * AND_MAX_HIST; e += blockDim.x) sG[e] = 0.0;
    // x = x0 ; history tiled with x0 (jaxopt init_state), residual history zero ; xin = a_col x^theta
    for (int64_t n = rb + tid; n < re; n += nth) {
        const double w0 = a.w_init[n];
        a.x[n] = w0;
        for (int i = 0; i < m; ++i) { a.X[i * a.ldv + n] = w0; a.R[i * a.ldv + n] = 0.0; }
        store_all_ranks(env, 0, n, op.stage_T(n, w0));
    }
    if (!all_sync(grid, env, epoch)) return;
    long long k = 0;
    double error = INFINITY;
    while (error > a.tol && k < a.max_iter) {
        const int pos = (int)(k % m);
        // f(x), residual, history update
        if (!op.apply_T(grid, env, epoch, sc, env.xin[env.rank][0], [&](int64_t n, double s) {
                const double y = 1.0 + beta * pow(s, inv_theta);
                const double xn = a.x[n];
                a.fx[n] = y;
                a.X[pos * a.ldv + n] = xn;
                a.R[pos * a.ldv + n] = y - xn;
            })) return;
        if (!all_sync(grid, env, epoch)) return;
        // new Gram row: <R_i, r>, i = 0..m-1 (8 per barrier)
        for (int i0 = 0; i0 < m; i0 += NVAL) {
            const int nb = (m - i0) < NVAL ? (m - i0) : NVAL;
            double hv[NVAL];
#pragma unroll
            for (int b = 0; b < NVAL; ++b) hv[b] = 0.0;
            const double *rcur = a.R + pos * a.ldv;
            for (int64_t n = rb + tid; n < re; n += nth) {
                const double rn = rcur[n];
#pragma unroll
                for (int b = 0; b < NVAL; ++b)
                    if (b < nb) hv[b] += a.R[(i0 + b) * a.ldv + n] * rn;
            }
            if (!grid_allreduce<NVAL, false>(grid, env, epoch, (i0 / NVAL) & 1 ? SET_X : SET_Y, hv, smem)) return;
            if (threadIdx.x == 0)
#pragma unroll
                for (int b = 0; b < NVAL; ++b)
                    if (b < nb) { sG[pos * AND_MAX_HIST + i0 + b] = hv[b]; sG[(i0 + b) * AND_MAX_HIST + pos] = hv[b]; }
        }
        __syncthreads();
        error = sqrt(sG[pos * AND_MAX_HIST + pos]);
        const bool extrapolate = (k >= m) && (k % a.mix == 0);
        if (extrapolate && threadIdx.x == 0) anderson_alphas(sG, m, a.ridge, sAlpha);
        __syncthreads();
        for (int64_t n = rb + tid; n < re; n += nth) {
            double xn;
            if (extrapolate) {
                double pa = 0.0, ra = 0.0;
                for (int i = 0; i < m; ++i) {
                    pa += sAlpha[i] * a.X[i * a.ldv + n];
                    ra += sAlpha[i] * a.R[i * a.ldv + n];
                }
                xn = pa + a.beta_mix * ra;
            } else {
                xn = a.fx[n];
            }
            a.x[n] = xn;
            store_all_ranks(env, 0, n, op.stage_T(n, xn));
        }
        if (!all_sync(grid, env, epoch)) return;
        ++k;
    }
    for (int64_t n = rb + tid; n < re; n += nth) a.w_out[n] = a.x[n];
    if (tid == 0) {
        env.status->iters = k;
        env.status->final_err = error;
        env.status->epoch_end = epoch;
    }
}

// ---------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Host side shared by the per-operator translation units (loops_dense.cu, loops_kron.cu,
// loops_cont.cu): each instantiates the three loop kernels for ONE operator type, so the
// three sets compile in parallel.
// ---------------------------------------------------------------------------
static int coop_grid(sdfs_ctx *ctx, const void *kern, int threads, size_t dyn_smem, int max_per_sm, int64_t work_groups, bool force_full, int *grid_out) {
    if (dyn_smem > 0) CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    int per_sm = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, dyn_smem));
    if (per_sm < 1) return sdfs_set_error(ctx, SDFS_ERR_CUDA, "cooperative kernel does not fit on an SM");
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    int64_t grid = (int64_t)per_sm * ctx->sm_count;
    if (!force_full) {   // small problems: fewer CTAs make the grid barrier cheaper
        int64_t want = work_groups < 1 ? 1 : work_groups;
        if (want < grid) grid = want;
    }
    if (grid > SDFS_MAX_GRID) grid = SDFS_MAX_GRID;
    *grid_out = (int)grid;
    return SDFS_OK;
}

static inline int64_t dense_groups(const DenseView &dv) {
    const int64_t g = (dv.row_end - dv.row_begin + TR - 1) / TR;
    return dv.vec2 ? g : (g + SDFS_WARPS - 1) / SDFS_WARPS;
}

enum { LOOP_SA = 0, LOOP_NEWTON = 1, LOOP_ANDERSON = 2 };

// `a` points to the SAArgs / NewtonArgs / AndersonArgs of the chosen loop
template <class Op, int WHICH>
static int loop_launch(sdfs_ctx *ctx, Op &lop, void *a, LoopEnv *env, size_t dyn_smem, int max_per_sm,
                       int64_t work_groups, bool force_full) {
    const void *kern;
    if constexpr (WHICH == LOOP_SA) kern = (const void *)k_sa_loop<Op>;
    else if constexpr (WHICH == LOOP_NEWTON) kern = (const void *)k_newton_loop<Op>;
    else kern = (const void *)k_anderson_loop<Op>;
    if constexpr (WHICH == LOOP_NEWTON) {      // GMRES Hessenberg workspace behind the operator's region
        if (((const NewtonArgs *)a)->krylov == SDFS_KRYLOV_GMRES) dyn_smem = ((dyn_smem + 15) & ~(size_t)15) + GMRES_WS_BYTES;
    }
    int grid = 0;
    TRY(coop_grid(ctx, kern, Op::kThreads, dyn_smem, max_per_sm, work_groups, force_full, &grid));
    void *args[] = {&lop, a, env};
    CUDA_TRY(ctx, cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(Op::kThreads), args, dyn_smem, ctx->stream));
    return SDFS_OK;
}

int loop_launch_dense(sdfs_op *op, int which, void *a, LoopEnv *env);   // loops_dense.cu
int loop_launch_kron(sdfs_op *op, int which, void *a, LoopEnv *env);    // loops_kron_{sa,newton,anderson}.cu
int loop_launch_cont(sdfs_op *op, int which, void *a, LoopEnv *env);    // loops_cont.cu
