// Operator handles and single applications: T, JVP, SDF, plain P x.
#include "common.cuh"
#include "rowdot.cuh"
#include "arena.cuh"
#include "cont.cuh"
#include "epilogue.cuh"
#include "kron_apply.cuh"

int factors_to_kron(const sdfs_factors *f, KronView *kv);
bool kron_can_shard(const KronView &kv);
void kron_restrict_leading(KronView *kv, int l0, int l1);
int launch_build_scalings(sdfs_ctx *ctx, const sdfs_factors *f, const KronView &kv, double gamma,
                          double theta, double mu_c, double *a_row, double *a_col, double *e_sdf);
int launch_expand_dense(sdfs_ctx *ctx, const KronView &kv, int64_t row_begin, int64_t row_end, int64_t ld, double *P);

#define TRY(x) do { int _rc = (x); if (_rc != SDFS_OK) return _rc; } while (0)

// ---------------------------------------------------------------------------
// Elementwise prologues (N work items; negligible next to the 8 N^2-byte stream)
// ---------------------------------------------------------------------------
// mode 0: x = a_col * w^theta                      (T prologue, ssy_wc_ratio.py:145)
// mode 1: x0 = a_col * w^theta, x1 = a_col * w^(theta-1) * v   (fused T + JVP)
// mode 2: x0 = a_col * w^theta, x1 = a_col * w^(theta-1)       (SDF)
// mode 3: x = v                                     (plain P x; stages into the aligned buffer)
template <bool FAST>   // FAST: exp(e log x) power form (factor-form path, where the N pows are a visible cost)
__global__ void k_prologue(int mode, int64_t N, const double *__restrict__ a_col, const double *__restrict__ w,
                           const double *__restrict__ v, double theta, double *__restrict__ x0,
                           double *__restrict__ x1) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (FAST && mode == 0) {        // two independent log/exp chains per thread (the T prologue at 10^7 states)
        for (; n + stride < N; n += 2 * stride) {
            const double wa = w[n], wb = w[n + stride];
            const double aa = a_col[n], ab = a_col[n + stride];
            const double ta = pow_pos(wa, theta), tb = pow_pos(wb, theta);
            x0[n] = aa * ta;
            x0[n + stride] = ab * tb;
        }
    }
    for (; n < N; n += stride) {
        if (mode == 3) { x0[n] = v[n]; continue; }
        const double wn = w[n];
        const double wt = FAST ? pow_pos(wn, theta) : pow(wn, theta);
        const double ac = a_col[n];
        x0[n] = ac * wt;
        if (mode == 1) x1[n] = ac * (wt / wn) * v[n];          // w^(theta-1) = w^theta / w
        else if (mode == 2) x1[n] = ac * (wt / wn);
    }
}

// elementwise epilogue at full occupancy (factor-form path: the contraction kernels run at 8-12
// warps per SM, far too few to hide the latency of one pow per output)
__global__ void k_epilogue_ew(int64_t N, const double *__restrict__ s0, const double *__restrict__ s1, EpiArgs e) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e.mode == 0) {              // T: two independent log/exp chains per thread
        for (; n + stride < N; n += 2 * stride) {
            const double xa = e.a_row[n] * s0[n], xb = e.a_row[n + stride] * s0[n + stride];
            const double pa = pow_pos(xa, e.inv_theta), pb = pow_pos(xb, e.inv_theta);
            e.out0[n] = 1.0 + e.beta * pa;
            e.out0[n + stride] = 1.0 + e.beta * pb;
        }
    }
    for (; n < N; n += stride) apply_epilogue<true>(e, n, s0[n], s1 ? s1[n] : s0[n]);
}

template <int NX>
__global__ void __launch_bounds__(SDFS_THREADS, 1)
k_dense_apply(const __grid_constant__ DenseView dv, const double *x0, const double *x1, EpiArgs e,
              const __grid_constant__ PeerArgs pa, DenseTail tail) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    RowPipe<NX> *rp = reinterpret_cast<RowPipe<NX> *>(dyn_smem);
    PipeState st;
    if (dv.vec2) pipe_init(rp, st);
    if (pa.nranks > 1) {
        dense_pass<NX>(dv, x0, x1, rp, st, [&](int64_t n, double s0, double s1) {
            const double val = epilogue_value(e, n, s0, s1);
            for (int r = 0; r < pa.nranks; ++r) pa.out[r][n] = val;
        }, tail);
        peer_exchange_finish(pa);
    } else {
        dense_pass<NX>(dv, x0, x1, rp, st, [&](int64_t n, double s0, double s1) { apply_epilogue<false>(e, n, s0, s1); }, tail);
    }
}

__global__ void __launch_bounds__(256) k_kron_mode(KronView kv, int m, const double *__restrict__ in, double *__restrict__ out) {
    extern __shared__ __align__(16) double kron_smem[];      // factor matrix + per-warp fragment stage
    kron_mode_apply_regpf(kv, m, in, kron_smem, KronSinkStore{out}, KronShare(kron_smem + KRON_SMAT_DOUBLES));
}

// 2-D TMA descriptor of the local row slice of P: dims (N columns, nloc rows), row pitch ld,
// box 256 columns x 8 rows, zero fill outside the matrix.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int dense_view_finalize(sdfs_ctx *ctx, DenseView *dv) {
    const int64_t nloc = dv->row_end - dv->row_begin;
    dv->vec2 = ((dv->ld % 2) == 0 && ((uintptr_t)dv->P % 16) == 0 && nloc > 0) ? 1 : 0;
    memset(&dv->tm, 0, sizeof(dv->tm));
    if (!dv->vec2) return SDFS_OK;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn) return sdfs_set_error(ctx, SDFS_ERR_UNSUPPORTED, "driver lacks cuTensorMapEncodeTiled");
        encode = (EncodeTiledFn)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)dv->N, (cuuint64_t)nloc};
    cuuint64_t strides[1] = {(cuuint64_t)dv->ld * 8};
    cuuint32_t box[2] = {TCW, TR};
    cuuint32_t es[2] = {1, 1};
    const CUresult r = encode(&dv->tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)dv->P, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dv->vec2 = 0;      // e.g. a pitch the TMA unit cannot address: use the direct-load path
        memset(&dv->tm, 0, sizeof(dv->tm));
    }
    return SDFS_OK;
}

// a_col along the axis of the first contraction (other coordinates 0)
__global__ void k_gather_strided(int n, long long stride, const double *__restrict__ src, double *__restrict__ dst) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[(long long)j * stride];
}
static int op_refresh_a_col_lead(sdfs_op *op) {
    sdfs_ctx *ctx = op->ctx;
    const KronMode &m0 = op->kv.modes[0];
    const int n = op->kv.shape[m0.dim];
    if (!op->a_col_lead) CUDA_TRY(ctx, cudaMalloc(&op->a_col_lead, (size_t)n * sizeof(double)));
    k_gather_strided<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, m0.stride, op->own_a_col, op->a_col_lead);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

// leading-axis slab of rank r: ceil(L / ranks) indices each, the last ranks possibly fewer or none
static inline void kron_slab(int L, int nranks, int r, int *l0, int *l1) {
    const int chunk = (L + nranks - 1) / nranks;
    int b = chunk * r, e = b + chunk;
    if (b > L) b = L;
    if (e > L) e = L;
    *l0 = b; *l1 = e;
}

// kvs = what this rank contracts: the whole view, or its slab restriction (call after every change of kv)
void op_sync_kvs(sdfs_op *op) {
    op->kvs = op->kv;
    if (op->kron_sharded) {
        int l0, l1;
        kron_slab(op->kv.shape[0], op->ctx->nranks, op->ctx->rank, &l0, &l1);
        kron_restrict_leading(&op->kvs, l0, l1);
    }
}

// rows of the result that rank r computes (dense row shards, factor-form slabs, or everything)
void op_rank_rows(const sdfs_op *op, int r, int64_t *rb, int64_t *re) {
    const int64_t N = op_N(op);
    const int G = op->ctx->nranks;
    if (op->storage == SDFS_STORAGE_KRON && op->kron_sharded) {
        int l0, l1;
        kron_slab(op->kv.shape[0], G, r, &l0, &l1);
        const int64_t inner = N / op->kv.shape[0];
        *rb = l0 * inner; *re = l1 * inner;
    } else if (op->storage == SDFS_STORAGE_DENSE && G > 1 && op->dv.row_end - op->dv.row_begin < N) {
        const int64_t chunk = (N + G - 1) / G;
        int64_t b = chunk * r, e = b + chunk;
        if (b > N) b = N;
        if (e > N) e = N;
        *rb = b; *re = e;
    } else {
        *rb = 0; *re = N;
    }
}
bool op_is_sharded(const sdfs_op *op) {
    int64_t rb, re;
    op_rank_rows(op, op->ctx->rank, &rb, &re);
    return op->ctx->nranks > 1 && re - rb < op_N(op);
}

int op_ensure_work(sdfs_op *op, int n_vectors) {
    sdfs_ctx *ctx = op->ctx;
    if (op->n_work >= n_vectors) return SDFS_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (op->work) CUDA_TRY(ctx, cudaFree(op->work));
    op->work = nullptr;
    const int64_t N = op_N(op);
    op->ldv = round_up(N, 64) + 64;
    CUDA_TRY(ctx, cudaMalloc(&op->work, (size_t)n_vectors * op->ldv * sizeof(double)));
    CUDA_TRY(ctx, cudaMemsetAsync(op->work, 0, (size_t)n_vectors * op->ldv * sizeof(double), ctx->stream));
    op->n_work = n_vectors;
    return SDFS_OK;
}

static inline int ew_grid(sdfs_ctx *ctx, int64_t N) {
    int64_t g = (N + 255) / 256;
    // several waves of CTAs: a grid of exactly 8 per SM leaves a quarter-occupancy tail wave when the
    // kernel's registers admit only 6 resident CTAs (k_epilogue_ew at 9.8 M states: 96 -> 7x us)
    const int64_t cap = (int64_t)ctx->sm_count * 32;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

static inline int dense_grid(sdfs_ctx *ctx, const DenseView &dv) {
    const int64_t nloc = dv.row_end - dv.row_begin;
    int64_t g = (nloc + TR - 1) / TR;                 // one CTA per SM sweeps groups of TR rows
    if (!dv.vec2) g = (g + SDFS_WARPS - 1) / SDFS_WARPS;
    const int64_t cap = ctx->sm_count;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// last partial wave of row groups split by columns over the whole grid (rowdot.cuh, DenseTail);
// SDFS_DENSE_TAIL=0 keeps whole groups (the A/B switch of profiles/r02_dense_tail.md)
static int op_dense_tail(sdfs_op *op, DenseTail *tail) {
    sdfs_ctx *ctx = op->ctx;
    static const bool tail_allowed = !(getenv("SDFS_DENSE_TAIL") && atoi(getenv("SDFS_DENSE_TAIL")) == 0);
    tail->buf = nullptr; tail->cnt = nullptr;
    if (!tail_allowed || !op->dv.vec2) return SDFS_OK;
    const size_t part_bytes = (size_t)ctx->sm_count * 2 * TR * sizeof(double);
    if (!op->dense_tail) {
        const size_t bytes = part_bytes + (size_t)ctx->sm_count * sizeof(unsigned int);
        CUDA_TRY(ctx, cudaMalloc(&op->dense_tail, bytes));
        CUDA_TRY(ctx, cudaMemsetAsync(op->dense_tail, 0, bytes, ctx->stream));
    }
    tail->buf = op->dense_tail;
    tail->cnt = (unsigned int *)((char *)op->dense_tail + part_bytes);
    return SDFS_OK;
}

template <int NX>
static int launch_dense_apply(sdfs_ctx *ctx, const DenseView &dv, const double *x0, const double *x1, const EpiArgs &e,
                              const PeerArgs &pa, DenseTail tail) {
    const size_t smem = dv.vec2 ? sizeof(RowPipe<NX>) : 0;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_dense_apply<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(RowPipe<NX>)));
    const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
    if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    k_dense_apply<NX><<<dense_grid(ctx, dv), SDFS_THREADS, smem, ctx->stream>>>(dv, x0, x1, e, pa, tail);
    if (prof) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

// Continuous-state operator: T w, and the fused T + J_T(w) v sweep (one loop over the shock nodes)
__global__ void __launch_bounds__(256) k_cont_T(ContView cv, const double *__restrict__ w, double *__restrict__ out) {
    const double inv_theta = 1.0 / cv.theta;
    cont_pass<0>(cv, w, nullptr, [&](int64_t n, double kg, double) { out[n] = 1.0 + cv.beta * pow(kg, inv_theta); });
}
__global__ void __launch_bounds__(256) k_cont_jvp(ContView cv, const double *__restrict__ w, const double *__restrict__ v,
                                                  double *__restrict__ out) {
    // J_T(w) v = beta s^((1-theta)/theta) const(x) L(v),  s = Kg(w)   (derivative of 1 + beta s^(1/theta))
    const double d_exp = (1.0 - cv.theta) / cv.theta;
    cont_pass<2>(cv, w, v, [&](int64_t n, double kg, double l) {
        out[n] = cv.beta * pow(kg, d_exp) * cont_rowfac(cv, n) * l;
    });
}

// lin_interp (utils.py:17-23) at M arbitrary points: x is D x M row-major
template <int D>
__global__ void k_interp_points(ContView cv, const double *__restrict__ vals, const double *__restrict__ x, int64_t M,
                                double *__restrict__ out) {
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        double xn[D], o0, o1;
#pragma unroll
        for (int d = 0; d < D; ++d) xn[d] = x[(int64_t)d * M + m];
        cont_interp<D, 1>(cv, xn, vals, vals, o0, o1);
        out[m] = o0;
    }
}

extern "C" int sdfs_interp_points(sdfs_ctx *ctx, int D, const int32_t *h_sizes, const double *h_g0, const double *h_intv,
                                  const double *d_vals, const double *d_x, int64_t M, double *d_out) {
    ARG_CHECK(ctx, ctx && h_sizes && h_g0 && h_intv && d_vals && d_x && d_out && M >= 0 && (D == 4 || D == 6));
    if (M == 0) return SDFS_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ContView cv;
    memset(&cv, 0, sizeof(cv));
    cv.D = D;
    for (int d = 0; d < D; ++d) {
        ARG_CHECK(ctx, h_sizes[d] >= 2 && h_intv[d] != 0.0);
        cv.n[d] = h_sizes[d]; cv.g0[d] = h_g0[d]; cv.intv[d] = h_intv[d];
    }
    const int grid = (int)((M + 255) / 256 < 4096 ? (M + 255) / 256 : 4096);
    if (D == 4) k_interp_points<4><<<grid, 256, 0, ctx->stream>>>(cv, d_vals, d_x, M, d_out);
    else k_interp_points<6><<<grid, 256, 0, ctx->stream>>>(cv, d_vals, d_x, M, d_out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

static int run_cont(sdfs_op *op, int which, const double *d_w, const double *d_v, double *d_out) {
    sdfs_ctx *ctx = op->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const ContView &cv = op->cv;
    int64_t g = (cv.N * 32 + 255) / 256;                 // one warp per state
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    const int grid = (int)(g < cap ? (g > 0 ? g : 1) : cap);
    if (which == 0) k_cont_T<<<grid, 256, 0, ctx->stream>>>(cv, d_w, d_out);
    else k_cont_jvp<<<grid, 256, 0, ctx->stream>>>(cv, d_w, d_v, d_out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

// One mode contraction out = (I x .. x M_m x .. x I) in of a factor-form view (also used by the
// batched sweep, sweep.cu, on a view with a leading column axis).
int launch_kron_mode(sdfs_ctx *ctx, const KronView &kv, int m, const double *in, double *out) {
    // launch shape of the tensor-core contraction; the two environment variables are tuning switches
    // (profiles/r01_kron_modes.md), clamped to shapes the kernel supports (__launch_bounds__(256))
    auto env_int = [](const char *name, int dflt, int lo, int hi) {
        const char *v = getenv(name);
        const int x = v ? atoi(v) : dflt;
        return x < lo || x > hi ? dflt : x;
    };
    static const int kron_tc_threads = env_int("SDFS_KRON_THREADS", 256, 32, 256) & ~31;
    static const int kron_tc_ctas = env_int("SDFS_KRON_CTAS", 2, 1, 8);
    const KronMode &md = kv.modes[m];
    // tensor-core contraction (n >= KRON_TC_MIN): 8 warps per CTA, each on tiles of 8 fibres, one contiguous
    // tile range per CTA; register-tiled FMA contraction (short axes): one fibre per thread,
    // work items dealt round-robin (any grid size is valid)
    const int nm = kv.shape[md.dim];
    const long long fibres = md.Fcount * md.Mcount;
    const bool tc = nm >= KRON_TC_MIN && nm <= KRON_NMAX_LIMIT;
    const int threads = tc ? kron_tc_threads : (fibres >= (long long)ctx->sm_count * 128 ? 128 : 64);
    const int grid = tc ? ctx->sm_count * kron_tc_ctas : ctx->sm_count * 6;
    static bool attr_set = false;
    if (!attr_set) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_kron_mode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KRON_APPLY_SMEM));
        attr_set = true;
    }
    k_kron_mode<<<grid, threads, KRON_APPLY_SMEM, ctx->stream>>>(kv, m, in, out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

// In-place all-gather of a full-length vector whose rows were computed by their owning ranks
int op_allgather(sdfs_op *op, double *d_vec) {
    int64_t rb[SDFS_MAX_RANKS], re[SDFS_MAX_RANKS];
    for (int r = 0; r < op->ctx->nranks; ++r) op_rank_rows(op, r, &rb[r], &re[r]);
    return comm_allgather_parts(op->ctx, d_vec, rb, re);
}

int launch_kron_apply(sdfs_ctx *ctx, const KronView &kv, const KronApplyArgs &ka, const EpiArgs &e, const PeerArgs &pa);   // kron_apply.cu

// Shared driver: prologue -> P pass(es) -> epilogue.
static int run_apply(sdfs_op *op, int pmode, const double *d_w, const double *d_v, EpiArgs e, bool gather0,
                     bool gather1) {
    sdfs_ctx *ctx = op->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool dense = op->storage == SDFS_STORAGE_DENSE;
    const int64_t N = dense ? op->dv.N : op->kv.N;
    e.inv_theta = 1.0 / e.theta;
    TRY(op_ensure_work(op, 4));
    double *x0 = op->work, *x1 = op->work + op->ldv;
    const int nx = (pmode == 1 || pmode == 2) ? 2 : 1;
    const double *a_col = dense ? op->dv.a_col : op->kv.a_col;
    const double theta = dense ? op->dv.theta : op->kv.theta;
    if (dense) {
        k_prologue<false><<<ew_grid(ctx, N), 256, 0, ctx->stream>>>(pmode, N, a_col, d_w, d_v, theta, x0, x1);
        ctx->launches++;
        const bool sharded = ctx->nranks > 1 && op->dv.row_end - op->dv.row_begin < N;
        PeerArgs pa;
        memset(&pa, 0, sizeof(pa));
        // fused exchange: single-output applies of a sharded operator once the arenas are mapped
        // (SDFS_FUSED_EXCHANGE=0 forces the NCCL all-gather path: used by the A/B measurement)
        static const bool fused_allowed = !(getenv("SDFS_FUSED_EXCHANGE") && atoi(getenv("SDFS_FUSED_EXCHANGE")) == 0);
        const bool fused = sharded && fused_allowed && e.mode != 2 && gather0 && e.out0 && comm_peers_ready(ctx) &&
                           comm_arena_maxN(ctx) >= N;
        if (fused) {
            unsigned long long *ep = comm_epoch(ctx);
            *ep += 1;                                           // every rank issues the same sequence
            pa.nranks = ctx->nranks; pa.rank = ctx->rank; pa.epoch = *ep;
            for (int r = 0; r < ctx->nranks; ++r) {
                void *base = comm_peer_arena(ctx, r);
                pa.out[r] = arena_apply_buf(base, comm_arena_maxN(ctx), (int)(*ep & 1));
                pa.sig[r] = (unsigned long long *)base + ctx->rank;
            }
            pa.mine = (unsigned long long *)comm_peer_arena(ctx, ctx->rank);
            pa.counter = (unsigned int *)((char *)ctx->d_status + 4096 - 64);
            pa.h_abort = ctx_h_abort(ctx);
        }
        // a rank without rows still takes part in the exchange (one CTA, empty pass)
        if (op->dv.row_end > op->dv.row_begin || fused) {
            DenseTail tail;
            TRY(op_dense_tail(op, &tail));
            if (nx == 1) TRY(launch_dense_apply<1>(ctx, op->dv, x0, x0, e, pa, tail));
            else TRY(launch_dense_apply<2>(ctx, op->dv, x0, x1, e, pa, tail));
        }
        CUDA_TRY(ctx, cudaGetLastError());
        if (fused) {
            CUDA_TRY(ctx, cudaMemcpyAsync(e.out0, pa.out[ctx->rank], (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        } else if (sharded) {
            if (gather0 && e.out0) TRY(op_allgather(op, e.out0));
            if (gather1 && e.out1) TRY(op_allgather(op, e.out1));
        }
    } else {
        const KronView &kv = op->kvs;
        if (!op->kron_tmp[0]) {
            CUDA_TRY(ctx, cudaMalloc(&op->kron_tmp[0], (size_t)N * sizeof(double)));
            CUDA_TRY(ctx, cudaMalloc(&op->kron_tmp[1], (size_t)N * sizeof(double)));
        }
        const bool sharded = op->kron_sharded;
        // Large whole operators: elementwise prologue / epilogue kernels at full occupancy around one launch per
        // mode.  The fused single launch wins while an application is launch- and barrier-bound (0.030 against
        // 0.039 ms at 10^5 states, 0.058 against 0.068 at 10^6); at 10^7 states the log/exp work dominates and
        // runs faster at 40+ warps per SM than inside the 16-warp contraction kernel.  SDFS_KRON_SPLIT_MIN
        // sets the switch-over (states; 0 = never split).
        static const long long split_min = getenv("SDFS_KRON_SPLIT_MIN") ? atoll(getenv("SDFS_KRON_SPLIT_MIN")) : (1LL << 21);
        if (!sharded && split_min > 0 && N >= split_min) {
            const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
            if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));     // whole application: prologue .. epilogue
            k_prologue<true><<<ew_grid(ctx, N), 256, 0, ctx->stream>>>(pmode, N, a_col, d_w, d_v, theta, x0, x1);
            ctx->launches++;
            double *sfin[2] = {op->work + 2 * op->ldv, op->work + 3 * op->ldv};
            for (int pass = 0; pass < nx; ++pass) {
                const double *in = (pass == 0) ? x0 : x1;
                for (int m = 0; m < kv.n_modes; ++m) {
                    double *out = (m == kv.n_modes - 1) ? sfin[pass] : op->kron_tmp[m & 1];
                    TRY(launch_kron_mode(ctx, kv, m, in, out));
                    in = out;
                }
            }
            k_epilogue_ew<<<ew_grid(ctx, N), 256, 0, ctx->stream>>>(N, sfin[0], nx > 1 ? sfin[1] : nullptr, e);
            ctx->launches++;
            if (prof) {
                CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
                ctx->prof_used += 2;
            }
            CUDA_TRY(ctx, cudaGetLastError());
            return SDFS_OK;
        }
        KronApplyArgs ka{};
        ka.pmode = pmode; ka.w = d_w; ka.v = d_v;
        ka.tmp0 = op->kron_tmp[0]; ka.tmp1 = op->kron_tmp[1];
        ka.s0 = op->work + 2 * op->ldv;
        PeerArgs pa;
        memset(&pa, 0, sizeof(pa));
        static const bool fused_allowed = !(getenv("SDFS_FUSED_EXCHANGE") && atoi(getenv("SDFS_FUSED_EXCHANGE")) == 0);
        const bool fused = sharded && fused_allowed && e.mode != 2 && gather0 && e.out0 && comm_peers_ready(ctx) &&
                           comm_arena_maxN(ctx) >= N;
        if (fused) {
            unsigned long long *ep = comm_epoch(ctx);
            *ep += 1;
            pa.nranks = ctx->nranks; pa.rank = ctx->rank; pa.epoch = *ep;
            for (int r = 0; r < ctx->nranks; ++r) {
                void *base = comm_peer_arena(ctx, r);
                pa.out[r] = arena_apply_buf(base, comm_arena_maxN(ctx), (int)(*ep & 1));
                pa.sig[r] = (unsigned long long *)base + ctx->rank;
            }
            pa.mine = (unsigned long long *)comm_peer_arena(ctx, ctx->rank);
            pa.counter = (unsigned int *)((char *)ctx->d_status + 4096 - 64);
            pa.h_abort = ctx_h_abort(ctx);
        }
        KronView kvl = kv;                                      // launch copy: a_col rides in the first factor matrix
        if (pmode != 3) kvl.modes[0].colscale = op->a_col_lead;
        TRY(launch_kron_apply(ctx, kvl, ka, e, pa));
        if (fused) {
            CUDA_TRY(ctx, cudaMemcpyAsync(e.out0, pa.out[ctx->rank], (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        } else if (sharded) {
            if (gather0 && e.out0) TRY(op_allgather(op, e.out0));
            if (gather1 && e.out1) TRY(op_allgather(op, e.out1));
        }
    }
    return SDFS_OK;
}

extern "C" {

int sdfs_op_from_dense(sdfs_ctx *ctx, const double *d_P, int64_t N, int64_t ld, int64_t row_begin,
                       int64_t row_end, const double *d_a_row, const double *d_a_col, double beta,
                       double theta, sdfs_op **out) {
    ARG_CHECK(ctx, ctx && out && d_P && d_a_row && d_a_col);
    ARG_CHECK(ctx, N >= 1 && ld >= N && row_begin >= 0 && row_begin <= row_end && row_end <= N);
    ARG_CHECK(ctx, theta != 0.0);
    sdfs_op *op = new sdfs_op();
    op->ctx = ctx;
    op->storage = SDFS_STORAGE_DENSE;
    op->dv.P = d_P; op->dv.N = N; op->dv.ld = ld;
    op->dv.row_begin = row_begin; op->dv.row_end = row_end;
    op->dv.a_row = d_a_row; op->dv.a_col = d_a_col; op->dv.e_sdf = nullptr;
    op->dv.beta = beta; op->dv.theta = theta;
    {
        const int rc = dense_view_finalize(ctx, &op->dv);
        if (rc) { delete op; return rc; }
    }
    *out = op;
    return SDFS_OK;
}

int sdfs_op_from_factors(sdfs_ctx *ctx, sdfs_factors *f, int storage, sdfs_op **out) {
    ARG_CHECK(ctx, ctx && f && out && f->ctx == ctx);
    ARG_CHECK(ctx, storage == SDFS_STORAGE_DENSE || storage == SDFS_STORAGE_KRON || storage == SDFS_STORAGE_DENSE_REPLICATED ||
                   storage == SDFS_STORAGE_KRON_LOCAL);
    const bool replicated = storage == SDFS_STORAGE_DENSE_REPLICATED;
    if (replicated) storage = SDFS_STORAGE_DENSE;
    const bool kron_local = storage == SDFS_STORAGE_KRON_LOCAL;
    if (kron_local) storage = SDFS_STORAGE_KRON;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    sdfs_op *op = new sdfs_op();
    op->ctx = ctx;
    op->storage = storage;
    op->factors = f;
    factors_to_kron(f, &op->kv);
    const int64_t N = op->kv.N;
    double beta, gamma, psi, mu_c;
    if (f->model == SDFS_MODEL_SSY) { beta = f->params[0]; gamma = f->params[1]; psi = f->params[2]; mu_c = f->params[3]; }
    else { beta = f->params[0]; psi = f->params[1]; gamma = f->params[2]; mu_c = f->params[5]; }
    op->gamma = gamma; op->psi = psi; op->mu_c = mu_c;
    const double theta = (1.0 - gamma) / (1.0 - 1.0 / psi);
    int rc = SDFS_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&op->own_a_row, N * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&op->own_a_col, N * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&op->own_e_sdf, N * sizeof(double))) != cudaSuccess) {
        rc = sdfs_set_error(ctx, SDFS_ERR_NOMEM, "operator scalings (N=%lld): %s", (long long)N, cudaGetErrorString(e));
        sdfs_op_destroy(op);
        return rc;
    }
    rc = launch_build_scalings(ctx, f, op->kv, gamma, theta, mu_c, op->own_a_row, op->own_a_col, op->own_e_sdf);
    if (rc) { sdfs_op_destroy(op); return rc; }
    op->kv.a_row = op->own_a_row; op->kv.a_col = op->own_a_col; op->kv.e_sdf = op->own_e_sdf;
    op->kv.beta = beta; op->kv.theta = theta;
    rc = op_refresh_a_col_lead(op);
    if (rc) { sdfs_op_destroy(op); return rc; }
    // slab-sharded factor form: leading axis split over the ranks (SDFS_KRON_SHARD=0 keeps every rank whole)
    // Below ~4 M states one GPU finishes an application in less time than the slab exchange costs (measured at
    // 8 ranks: (32,)^4 0.100 ms sharded against 0.058 ms whole; (56,)^4 0.32 against 0.40 ms), so smaller
    // operators stay whole on every rank.  SDFS_KRON_SHARD_MIN overrides the threshold (0 = always shard).
    static const bool shard_allowed = !(getenv("SDFS_KRON_SHARD") && atoi(getenv("SDFS_KRON_SHARD")) == 0);
    static const long long shard_min = getenv("SDFS_KRON_SHARD_MIN") ? atoll(getenv("SDFS_KRON_SHARD_MIN")) : (1LL << 22);
    op->kron_sharded = storage == SDFS_STORAGE_KRON && !kron_local && shard_allowed && ctx->nranks > 1 &&
                       kron_can_shard(op->kv) && op->kv.shape[0] >= ctx->nranks && N >= shard_min;
    op_sync_kvs(op);
    if (storage == SDFS_STORAGE_DENSE) {
        const int64_t ld = round_up(N, 64);
        const int64_t chunk = (N + ctx->nranks - 1) / ctx->nranks;
        int64_t rb = chunk * ctx->rank, re = rb + chunk;
        if (rb > N) rb = N;
        if (re > N) re = N;
        if (replicated) { rb = 0; re = N; }
        const size_t bytes = (size_t)(re - rb > 0 ? re - rb : 1) * ld * sizeof(double);
        if ((e = cudaMalloc(&op->own_P, bytes)) != cudaSuccess) {
            rc = sdfs_set_error(ctx, SDFS_ERR_NOMEM,
                                "dense P needs %.2f GB on this rank (N=%lld, rows %lld..%lld): %s; use "
                                "SDFS_STORAGE_KRON or more ranks", bytes / 1e9, (long long)N, (long long)rb,
                                (long long)re, cudaGetErrorString(e));
            sdfs_op_destroy(op);
            return rc;
        }
        rc = launch_expand_dense(ctx, op->kv, rb, re, ld, op->own_P);
        if (rc) { sdfs_op_destroy(op); return rc; }
        op->dv.P = op->own_P; op->dv.N = N; op->dv.ld = ld;
        op->dv.row_begin = rb; op->dv.row_end = re;
        op->dv.a_row = op->own_a_row; op->dv.a_col = op->own_a_col; op->dv.e_sdf = op->own_e_sdf;
        op->dv.beta = beta; op->dv.theta = theta;
        rc = dense_view_finalize(ctx, &op->dv);
        if (rc) { sdfs_op_destroy(op); return rc; }
    }
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        rc = sdfs_set_error(ctx, SDFS_ERR_CUDA, "operator build: %s", cudaGetErrorString(e));
        sdfs_op_destroy(op);
        return rc;
    }
    *out = op;
    return SDFS_OK;
}

// Continuous-state operator: h_grids = the D uniform grids concatenated (sizes h_sizes), h_nodes =
// D x Q shocks (row d = shocks of state component d), h_weights = Q weights.
int sdfs_op_continuous(sdfs_ctx *ctx, int model, const double *h_params, const int32_t *h_sizes,
                       const double *h_grids, const double *h_nodes, const double *h_weights, int64_t Q,
                       sdfs_op **out) {
    ARG_CHECK(ctx, ctx && h_params && h_sizes && h_grids && h_nodes && h_weights && out && Q >= 1 && Q < (1ll << 30));
    ARG_CHECK(ctx, model == SDFS_MODEL_SSY || model == SDFS_MODEL_GCY);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int D = (model == SDFS_MODEL_SSY) ? 4 : 6;
    int64_t N = 1, gtot = 0;
    for (int d = 0; d < D; ++d) {
        if (h_sizes[d] < 2) return sdfs_set_error(ctx, SDFS_ERR_ARG, "grid %d needs at least 2 points", d);
        N *= h_sizes[d];
        gtot += h_sizes[d];
    }
    sdfs_op *op = new sdfs_op();
    op->ctx = ctx;
    op->storage = SDFS_STORAGE_CONT;
    const size_t doubles = (size_t)gtot + (size_t)D * Q + (size_t)Q;
    cudaError_t e = cudaMalloc(&op->cont_mem, doubles * sizeof(double));
    if (e != cudaSuccess) { delete op; return sdfs_set_error(ctx, SDFS_ERR_NOMEM, "continuous operator tables: %s", cudaGetErrorString(e)); }
    ContView &cv = op->cv;
    memset(&cv, 0, sizeof(cv));
    cv.model = model; cv.D = D; cv.Q = (int)Q; cv.N = N;
    double *dptr = op->cont_mem;
    const double *hg = h_grids;
    cudaError_t up = cudaSuccess;              // first failing upload (checked once, below: the op must be destroyed)
    auto upload = [&](double *dst, const double *src, size_t count) {
        const cudaError_t r = cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        if (up == cudaSuccess) up = r;
    };
    for (int d = 0; d < D; ++d) {
        cv.n[d] = h_sizes[d];
        cv.grid[d] = dptr;
        cv.g0[d] = hg[0];
        cv.intv[d] = hg[1] - hg[0];
        upload(dptr, hg, (size_t)h_sizes[d]);
        dptr += h_sizes[d];
        hg += h_sizes[d];
    }
    cv.nodes = dptr;
    upload(dptr, h_nodes, (size_t)D * Q);
    dptr += (size_t)D * Q;
    cv.weights = dptr;
    upload(dptr, h_weights, (size_t)Q);
    const int np = (model == SDFS_MODEL_SSY) ? 13 : 18;
    for (int i = 0; i < np; ++i) cv.p[i] = h_params[i];
    double psi;
    if (model == SDFS_MODEL_SSY) { cv.beta = h_params[0]; cv.gamma = h_params[1]; psi = h_params[2]; cv.mu_c = h_params[3]; cv.phi_c = h_params[6]; }
    else { cv.beta = h_params[0]; psi = h_params[1]; cv.gamma = h_params[2]; cv.mu_c = h_params[5]; cv.phi_c = h_params[6]; }
    cv.theta = (1.0 - cv.gamma) / (1.0 - 1.0 / psi);
    cv.row_begin = 0;
    cv.row_end = N;      // states are not sharded: every rank evaluates the whole grid
    e = cudaStreamSynchronize(ctx->stream);
    if (up != cudaSuccess) e = up;
    if (e != cudaSuccess) { (void)cudaGetLastError(); sdfs_op_destroy(op); return sdfs_set_error(ctx, SDFS_ERR_CUDA, "continuous operator upload: %s", cudaGetErrorString(e)); }
    *out = op;
    return SDFS_OK;
}

int sdfs_op_destroy(sdfs_op *op) {
    if (!op) return SDFS_OK;
    sdfs_ctx *ctx = op->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (op->cont_mem) cudaFree(op->cont_mem);
    if (op->own_P) cudaFree(op->own_P);
    if (op->own_a_row) cudaFree(op->own_a_row);
    if (op->own_a_col) cudaFree(op->own_a_col);
    if (op->own_e_sdf) cudaFree(op->own_e_sdf);
    if (op->work) cudaFree(op->work);
    if (op->slots) cudaFree(op->slots);
    if (op->dense_tail) cudaFree(op->dense_tail);
    if (op->kron_tmp[0]) cudaFree(op->kron_tmp[0]);
    if (op->kron_tmp[1]) cudaFree(op->kron_tmp[1]);
    if (op->a_col_lead) cudaFree(op->a_col_lead);
    delete op;
    return SDFS_OK;
}

int sdfs_op_info(sdfs_op *op, int64_t *N, int64_t *ld, int64_t *row_begin, int64_t *row_end, double *beta,
                 double *theta, int *storage) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_info: NULL op");
    const bool dense = op->storage == SDFS_STORAGE_DENSE;
    const bool cont = op->storage == SDFS_STORAGE_CONT;
    if (N) *N = op_N(op);
    if (ld) *ld = dense ? op->dv.ld : 0;
    int64_t rb = 0, re = op_N(op);
    if (dense) { rb = op->dv.row_begin; re = op->dv.row_end; }
    else if (op->storage == SDFS_STORAGE_KRON) { rb = op->kvs.row_begin; re = op->kvs.row_end; }
    if (row_begin) *row_begin = rb;
    if (row_end) *row_end = re;
    if (beta) *beta = dense ? op->dv.beta : (cont ? op->cv.beta : op->kv.beta);
    if (theta) *theta = dense ? op->dv.theta : (cont ? op->cv.theta : op->kv.theta);
    if (storage) *storage = op->storage;
    return SDFS_OK;
}

int sdfs_op_arrays(sdfs_op *op, const double **d_P, const double **d_a_row, const double **d_a_col,
                   const double **d_e_sdf) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_arrays: NULL op");
    const bool dense = op->storage == SDFS_STORAGE_DENSE;
    if (d_P) *d_P = dense ? op->dv.P : nullptr;
    if (d_a_row) *d_a_row = dense ? op->dv.a_row : op->kv.a_row;
    if (d_a_col) *d_a_col = dense ? op->dv.a_col : op->kv.a_col;
    if (d_e_sdf) *d_e_sdf = dense ? op->dv.e_sdf : op->kv.e_sdf;
    return SDFS_OK;
}

int sdfs_op_set_esdf(sdfs_op *op, const double *d_e_sdf) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_set_esdf: NULL op");
    op->dv.e_sdf = d_e_sdf;
    op->kv.e_sdf = d_e_sdf;
    return SDFS_OK;
}

int sdfs_op_set_preferences(sdfs_op *op, double gamma, double psi, double beta) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_set_preferences: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, op->factors != nullptr && op->own_a_row != nullptr);
    ARG_CHECK(ctx, psi != 1.0 && gamma != 1.0);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const double theta = (1.0 - gamma) / (1.0 - 1.0 / psi);
    TRY(launch_build_scalings(ctx, op->factors, op->kv, gamma, theta, op->mu_c, op->own_a_row, op->own_a_col,
                              op->own_e_sdf));
    TRY(op_refresh_a_col_lead(op));
    op->gamma = gamma; op->psi = psi;
    op->kv.beta = beta; op->kv.theta = theta;
    op->dv.beta = beta; op->dv.theta = theta;
    op_sync_kvs(op);
    return SDFS_OK;
}

int sdfs_op_apply_T(sdfs_op *op, const double *d_w_in, double *d_w_out) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_apply_T: NULL op");
    ARG_CHECK(op->ctx, d_w_in && d_w_out);
    if (op->storage == SDFS_STORAGE_CONT) return run_cont(op, 0, d_w_in, nullptr, d_w_out);
    const bool dense = op->storage == SDFS_STORAGE_DENSE;
    EpiArgs e{0, dense ? op->dv.a_row : op->kv.a_row, nullptr, nullptr, dense ? op->dv.beta : op->kv.beta,
              dense ? op->dv.theta : op->kv.theta, d_w_out, nullptr};
    return run_apply(op, 0, d_w_in, nullptr, e, true, false);
}

int sdfs_op_apply_jvp(sdfs_op *op, const double *d_w, const double *d_v, double *d_out) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_apply_jvp: NULL op");
    ARG_CHECK(op->ctx, d_w && d_v && d_out);
    if (op->storage == SDFS_STORAGE_CONT) return run_cont(op, 1, d_w, d_v, d_out);
    const bool dense = op->storage == SDFS_STORAGE_DENSE;
    EpiArgs e{1, dense ? op->dv.a_row : op->kv.a_row, nullptr, nullptr, dense ? op->dv.beta : op->kv.beta,
              dense ? op->dv.theta : op->kv.theta, d_out, nullptr};
    return run_apply(op, 1, d_w, d_v, e, true, false);
}

int sdfs_op_apply_P(sdfs_op *op, const double *d_x, double *d_y) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_apply_P: NULL op");
    ARG_CHECK(op->ctx, d_x && d_y);
    if (op->storage == SDFS_STORAGE_CONT)
        return sdfs_set_error(op->ctx, SDFS_ERR_UNSUPPORTED, "continuous-state operators have no transition matrix P");
    EpiArgs e{3, nullptr, nullptr, nullptr, 0.0, 1.0, d_y, nullptr};
    return run_apply(op, 3, nullptr, d_x, e, true, false);
}

// Back-to-back launches of the dense pass alone (no prologue, no allocation, no host work in
// between): the sustained-rate measurement behind profiles/ and a diagnostic for the bench.
// mode 0 = T epilogue, 3 = plain P x.  x must have been staged by a previous apply.
int sdfs_op_bench_pass(sdfs_op *op, int mode, int reps, double *avg_ms) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_bench_pass: NULL op");
    sdfs_ctx *ctx = op->ctx;
    ARG_CHECK(ctx, op->storage == SDFS_STORAGE_DENSE && reps >= 1 && avg_ms && (mode == 0 || mode == 3));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(op_ensure_work(op, 4));
    double *x0 = op->work, *scratch = op->work + 3 * op->ldv;
    EpiArgs e{mode, op->dv.a_row, nullptr, nullptr, op->dv.beta, op->dv.theta, scratch, nullptr};
    e.inv_theta = 1.0 / e.theta;
    PeerArgs pa;
    memset(&pa, 0, sizeof(pa));                 // local pass only
    DenseTail tail;
    TRY(op_dense_tail(op, &tail));
    TRY(launch_dense_apply<1>(ctx, op->dv, x0, x0, e, pa, tail));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < reps; ++i) TRY(launch_dense_apply<1>(ctx, op->dv, x0, x0, e, pa, tail));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *avg_ms = (double)ms / reps;
    return SDFS_OK;
}

int sdfs_op_sdf(sdfs_op *op, const double *d_w, double *d_qf, double *d_euler) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_sdf: NULL op");
    if (op->storage == SDFS_STORAGE_CONT)
        return sdfs_set_error(op->ctx, SDFS_ERR_UNSUPPORTED, "SDF evaluation is defined for the discretised Markov-grid operators");
    const bool dense = op->storage == SDFS_STORAGE_DENSE;
    const double *es = dense ? op->dv.e_sdf : op->kv.e_sdf;
    ARG_CHECK(op->ctx, d_w && (d_qf || d_euler));
    if (d_qf && !es)
        return sdfs_set_error(op->ctx, SDFS_ERR_ARG, "sdfs_op_sdf: operator has no e_sdf vector (sdfs_op_set_esdf)");
    EpiArgs e{2, dense ? op->dv.a_row : op->kv.a_row, es, d_w, dense ? op->dv.beta : op->kv.beta,
              dense ? op->dv.theta : op->kv.theta, d_qf, d_euler};
    return run_apply(op, 2, d_w, nullptr, e, true, true);
}

}  // extern "C"

// Explicit rows of the SDF matrix Mbar(n, n')
__global__ void k_sdf_rows(int64_t N, const double *__restrict__ w, const double *__restrict__ a_col,
                           const double *__restrict__ e_sdf, double beta, double theta,
                           const int64_t *__restrict__ rows, double *__restrict__ out) {
    const int64_t n = rows[blockIdx.y];
    const double bt = pow(beta, theta) * e_sdf[n];
    const double den = w[n] - 1.0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < N; c += (int64_t)gridDim.x * blockDim.x)
        out[(int64_t)blockIdx.y * N + c] = bt * a_col[c] * pow(w[c] / den, theta - 1.0);
}

extern "C" int sdfs_op_sdf_rows(sdfs_op *op, const double *d_w, const int64_t *h_rows, int64_t n_rows,
                                double *d_out) {
    if (!op) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_op_sdf_rows: NULL op");
    sdfs_ctx *ctx = op->ctx;
    const bool dense = op->storage == SDFS_STORAGE_DENSE;
    const int64_t N = dense ? op->dv.N : op->kv.N;
    const double *es = dense ? op->dv.e_sdf : op->kv.e_sdf;
    ARG_CHECK(ctx, d_w && h_rows && d_out && n_rows >= 1 && n_rows <= 65535 && es);
    for (int64_t i = 0; i < n_rows; ++i) ARG_CHECK(ctx, h_rows[i] >= 0 && h_rows[i] < N);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int64_t *d_rows = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_rows, n_rows * sizeof(int64_t)));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_rows, h_rows, n_rows * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((unsigned)((N + 255) / 256 < 1024 ? (N + 255) / 256 : 1024), (unsigned)n_rows);
    k_sdf_rows<<<grid, 256, 0, ctx->stream>>>(N, d_w, dense ? op->dv.a_col : op->kv.a_col, es,
                                              dense ? op->dv.beta : op->kv.beta,
                                              dense ? op->dv.theta : op->kv.theta, d_rows, d_out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaFree(d_rows));
    return SDFS_OK;
}
