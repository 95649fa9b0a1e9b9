// The fused factor-form apply kernel in its own translation unit (its tensor-core contraction variants are
// the slowest code of the library to compile; the build runs one nvcc per unit in parallel).
#include "common.cuh"
#include "rowdot.cuh"
#include "epilogue.cuh"
#include "kron_apply.cuh"

// One application of a factor-form operator in one cooperative launch (kron_apply.cuh): prologue fused
// into the first contraction's loads, epilogue into the last contraction's sink, result rows stored to every
// rank when the view is slab-sharded.
struct KronEpi {
    EpiArgs e;
    PeerArgs pa;
    __device__ __forceinline__ void put(int64_t n, double val) const {
        if (pa.nranks > 1) { for (int r = 0; r < pa.nranks; ++r) pa.out[r][n] = val; }
        else e.out0[n] = val;
    }
    // per-row factor of the contraction (1 for the plain P x epilogue)
    __device__ __forceinline__ double pre(int64_t n) const { return e.mode == 3 ? 1.0 : e.a_row[n]; }
    // as0 = a_row s0, as1 = a_row s1 (same value when one contraction feeds the epilogue)
    __device__ __forceinline__ void one(int64_t n, double as0, double as1) const {
        if (e.mode == 0) put(n, 1.0 + e.beta * pow_pos(as0, e.inv_theta));
        else if (e.mode == 1) put(n, e.beta * pow_pos(as0, e.inv_theta - 1.0) * as1);
        else if (e.mode == 2) {          // two outputs: local stores (gathered by the host path)
            const double bt = pow(e.beta, e.theta);
            const double wm1 = e.w[n] - 1.0;
            if (e.out0) e.out0[n] = bt * e.e_sdf[n] * pow_pos(wm1, 1.0 - e.theta) * (as1 / e.a_row[n]);
            if (e.out1) e.out1[n] = bt * as0 / pow_pos(wm1, e.theta) - 1.0;
        } else put(n, as0);
    }
    __device__ __forceinline__ void four(int64_t n0, int64_t n1, int64_t n2, int64_t n3, double a, double b, double c, double d) const {
        if (e.mode == 0) {                  // T: four independent log/exp chains
            const double p0 = pow_pos(a, e.inv_theta), p1 = pow_pos(b, e.inv_theta);
            const double p2 = pow_pos(c, e.inv_theta), p3 = pow_pos(d, e.inv_theta);
            put(n0, 1.0 + e.beta * p0); put(n1, 1.0 + e.beta * p1);
            put(n2, 1.0 + e.beta * p2); put(n3, 1.0 + e.beta * p3);
        } else {
            one(n0, a, a); one(n1, b, b); one(n2, c, c); one(n3, d, d);
        }
    }
};

__global__ void __launch_bounds__(KRON_APPLY_THREADS, 2)
k_kron_apply(const __grid_constant__ KronView kv, const __grid_constant__ KronApplyArgs a, const __grid_constant__ KronEpi epi) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) double kron_smem[];
    kron_apply_device(grid, kv, a, kron_smem, kron_smem + KRON_SMAT_DOUBLES, epi);
    if (epi.pa.nranks > 1) peer_exchange_finish(epi.pa);
}

int launch_kron_apply(sdfs_ctx *ctx, const KronView &kv, const KronApplyArgs &ka, const EpiArgs &e, const PeerArgs &pa) {
    static int per_sm = 0;
    if (!per_sm) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_kron_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KRON_APPLY_SMEM));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_kron_apply, KRON_APPLY_THREADS, KRON_APPLY_SMEM));
        if (per_sm < 1) return sdfs_set_error(ctx, SDFS_ERR_CUDA, "k_kron_apply does not fit on an SM");
        if (per_sm > 2) per_sm = 2;
    }
    if (kv.n_modes < 2) return sdfs_set_error(ctx, SDFS_ERR_UNSUPPORTED, "factor-form apply needs at least two modes");
    // persistent grid: every CTA takes one contiguous range of fibre tiles per mode; small grids get fewer
    // CTAs (a warp per 8-fibre tile of the largest mode is all the parallelism there is)
    long long tiles = 1;
    for (int m = 0; m < kv.n_modes; ++m) {
        const KronMode &md = kv.modes[m];
        const long long t = md.Mcount * ((md.Fcount + 7) / 8);
        if (t > tiles) tiles = t;
    }
    long long grid = (tiles + (KRON_APPLY_THREADS / 32) - 1) / (KRON_APPLY_THREADS / 32);
    const long long cap = (long long)per_sm * ctx->sm_count;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    KronEpi epi{e, pa};
    // SDFS_KRON_TRACE=1: device timestamps at the phase boundaries of every apply, printed to stderr (diagnostic)
    static const bool trace_on = getenv("SDFS_KRON_TRACE") && atoi(getenv("SDFS_KRON_TRACE")) != 0;
    static unsigned long long *d_trace = nullptr;
    KronApplyArgs ka2 = ka;
    if (trace_on) {
        if (!d_trace) CUDA_TRY(ctx, cudaMalloc(&d_trace, 32 * sizeof(unsigned long long)));
        CUDA_TRY(ctx, cudaMemsetAsync(d_trace, 0, 32 * sizeof(unsigned long long), ctx->stream));
        ka2.trace = d_trace;
    }
    void *args[] = {(void *)&kv, (void *)&ka2, (void *)&epi};
    const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
    if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    CUDA_TRY(ctx, cudaLaunchCooperativeKernel((const void *)k_kron_apply, dim3((unsigned)grid), dim3(KRON_APPLY_THREADS), args,
                                              KRON_APPLY_SMEM, ctx->stream));
    if (prof) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    ctx->launches++;
    if (trace_on) {
        unsigned long long h[32];
        CUDA_TRY(ctx, cudaMemcpyAsync(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "kron_apply trace (us per phase, grid %lld):", grid);
        for (int i = 1; i < 32 && h[i]; ++i) fprintf(stderr, " %.1f", (h[i] - h[i - 1]) * 1e-3);
        fprintf(stderr, "\n");
    }
    return SDFS_OK;
}

