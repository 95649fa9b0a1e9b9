// Exchange arena of one rank (CUDA-IPC mapped by every peer) and the system-scope flag helpers.
// Layout: [flags 1 KB][reduction slots][x0][x1][y0][y1]
//   flags[src]   arrival counter written by rank src (monotone barrier epoch, shared by the loop
//                kernels' all_sync and the fused single-application exchange)
//   x0, x1       matvec inputs of the persistent loop kernels
//   y0, y1       results of fused single applications (double-buffered by epoch parity)
#pragma once
#include "common.cuh"

#define NSETS 8            // reduction slot sets (one per phase, see the phase tables in loops.cuh)
#define NVAL 8             // values per set (GMRES orthogonalises against 8 basis vectors per barrier)

static inline size_t arena_slots_doubles() { return (size_t)NSETS * SDFS_MAX_RANKS * SDFS_MAX_GRID * NVAL; }
static inline size_t arena_ldv(int64_t maxN) { return (size_t)round_up(maxN, 64) + 64; }
static inline size_t arena_bytes_for(int64_t maxN) {
    return 1024 + arena_slots_doubles() * sizeof(double) + 4 * arena_ldv(maxN) * sizeof(double);
}
static inline void arena_carve(void *base, int64_t maxN, unsigned long long **flags, double **slots, double **x0, double **x1) {
    char *b = (char *)base;
    *flags = (unsigned long long *)b;
    *slots = (double *)(b + 1024);
    *x0 = *slots + arena_slots_doubles();
    *x1 = *x0 + arena_ldv(maxN);
}
static inline double *arena_apply_buf(void *base, int64_t maxN, int which) {
    double *x0 = (double *)((char *)base + 1024) + arena_slots_doubles();
    return x0 + (2 + which) * arena_ldv(maxN);
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
#define SDFS_PEER_TIMEOUT_CLOCKS 60000000000LL   // ~30 s: a peer died

// accessors (comm.cu)
bool comm_peers_ready(sdfs_ctx *ctx);
void *comm_peer_arena(sdfs_ctx *ctx, int r);
int64_t comm_arena_maxN(sdfs_ctx *ctx);
unsigned long long *comm_epoch(sdfs_ctx *ctx);
int comm_allgather_rows(sdfs_ctx *ctx, double *d_vec, int64_t N);
int comm_allgather_parts(sdfs_ctx *ctx, double *d_vec, const int64_t *rb, const int64_t *re);
