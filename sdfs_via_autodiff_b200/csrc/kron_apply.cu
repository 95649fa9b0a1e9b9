// The fused factor-form apply kernel in its own translation unit (its tensor-core contraction variants are
// the slowest code of the library to compile; the build runs one nvcc per unit in parallel).
#ifndef KRON_APPLY_CTAS
#define KRON_APPLY_CTAS 2         // CTAs per SM of the fused apply (3 = 256 threads at <= 85 registers with a single-buffered stage: measured no gain)
#endif
#if KRON_APPLY_CTAS >= 3
#define KRON_STAGE_BUFS 1
#endif
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
    __device__ __forceinline__ bool single_output() const { return e.mode != 2; }
    __device__ __forceinline__ void put2(int64_t n, double v0, double v1) const {      // n even: one 16-byte store per rank
        const double2 v = make_double2(v0, v1);
        if (pa.nranks > 1) { for (int r = 0; r < pa.nranks; ++r) *reinterpret_cast<double2 *>(pa.out[r] + n) = v; }
        else *reinterpret_cast<double2 *>(e.out0 + n) = v;
    }
    // four single-output epilogues at once (independent log/exp chains): (as0, as1) pairs in, results in as1
    __device__ __forceinline__ void values4(double a0, double &a, double b0, double &b, double c0, double &c, double d0, double &d) const {
        if (e.mode == 0) {
            const double p0 = pow_pos(a0, e.inv_theta), p1 = pow_pos(b0, e.inv_theta);
            const double p2 = pow_pos(c0, e.inv_theta), p3 = pow_pos(d0, e.inv_theta);
            a = 1.0 + e.beta * p0; b = 1.0 + e.beta * p1; c = 1.0 + e.beta * p2; d = 1.0 + e.beta * p3;
        } else if (e.mode == 1) {
            const double ex = e.inv_theta - 1.0;
            const double p0 = pow_pos(a0, ex), p1 = pow_pos(b0, ex), p2 = pow_pos(c0, ex), p3 = pow_pos(d0, ex);
            a = e.beta * p0 * a; b = e.beta * p1 * b; c = e.beta * p2 * c; d = e.beta * p3 * d;
        }                                   // mode 3 (plain P x): the values are the contractions themselves
    }
};

__global__ void __launch_bounds__(KRON_APPLY_THREADS, KRON_APPLY_CTAS)
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
        if (per_sm > KRON_APPLY_CTAS) per_sm = KRON_APPLY_CTAS;
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

