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
    __device__ __forceinline__ void operator()(int64_t n, double s0, double s1) const {
        if (e.mode == 2) apply_epilogue<true>(e, n, s0, s1);            // two outputs: local stores (gathered by the host path)
        else put(n, epilogue_value<true>(e, n, s0, s1));
    }
    __device__ __forceinline__ void pair(int64_t n0, int64_t n1, double s0a, double s1a, double s0b, double s1b) const {
        if (e.mode == 0) {                  // T: two independent log/exp chains
            const double xa = e.a_row[n0] * s0a, xb = e.a_row[n1] * s0b;
            const double pa_ = pow_pos(xa, e.inv_theta), pb_ = pow_pos(xb, e.inv_theta);
            put(n0, 1.0 + e.beta * pa_);
            put(n1, 1.0 + e.beta * pb_);
        } else {
            (*this)(n0, s0a, s1a);
            (*this)(n1, s0b, s1b);
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
    void *args[] = {(void *)&kv, (void *)&ka, (void *)&epi};
    const bool prof = ctx->prof_on && ctx->prof_used + 2 <= ctx->prof_ev.size();
    if (prof) CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used], ctx->stream));
    CUDA_TRY(ctx, cudaLaunchCooperativeKernel((const void *)k_kron_apply, dim3((unsigned)grid), dim3(KRON_APPLY_THREADS), args,
                                              KRON_APPLY_SMEM, ctx->stream));
    if (prof) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], ctx->stream));
        ctx->prof_used += 2;
    }
    ctx->launches++;
    return SDFS_OK;
}

