// Multi-GPU plumbing: one process per GPU.  NCCL (loaded at run time from the
// wheel that ships with the image) for stream-ordered all-gathers of row slices,
// and a CUDA-IPC mapped exchange arena for the fused solver loops (peer stores
// over NVLink + flag barrier, loops.cu).
#include "common.cuh"
#include "arena.cuh"
#include <dlfcn.h>
#include <stdlib.h>

// Minimal NCCL ABI (nccl.h 2.2x): only what is called here.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8 };   // ncclDataType_t: int8 0,uint8 1,int32 2,uint32 3,int64 4,uint64 5,f16 6,f32 7,f64 8
enum { ncclSumOp = 0 };

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(sdfs_ctx *ctx) {
    if (g_nccl.lib) return SDFS_OK;
    // SDFS_NCCL_LIB (the Python package points it at the NCCL wheel it finds, _lib.py), else a libnccl this
    // process has already loaded (torch), else the loader path - no image-specific path is compiled in
    void *h = nullptr;
    if (const char *envp = getenv("SDFS_NCCL_LIB")) h = dlopen(envp, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return sdfs_set_error(ctx, SDFS_ERR_COMM, "cannot load libnccl.so.2 (set SDFS_NCCL_LIB): %s", dlerror());
#define SYM(field, name)                                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, name);                                              \
    if (!g_nccl.field) return sdfs_set_error(ctx, SDFS_ERR_COMM, "libnccl: missing symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(Broadcast, "ncclBroadcast");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = h;
    return SDFS_OK;
}

struct sdfs_comm_state {
    ncclComm_t comm = nullptr;
    // exchange arena
    void *arena = nullptr;               // own allocation
    size_t arena_bytes = 0;
    int64_t arena_maxN = 0;
    void *peer_arena[SDFS_MAX_RANKS] = {nullptr};   // mapped peers (own entry = arena)
    bool peers_mapped = false;
    unsigned long long epoch = 0;        // barrier epoch, identical on all ranks
};

#define NCCL_TRY(ctx, expr)                                                                      \
    do {                                                                                         \
        ncclResult_t _r = (expr);                                                                \
        if (_r != 0)                                                                             \
            return sdfs_set_error((ctx), SDFS_ERR_COMM, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, \
                                  g_nccl.GetErrorString(_r));                                    \
    } while (0)

int comm_destroy(sdfs_ctx *ctx) {
    if (!ctx || !ctx->comm) return SDFS_OK;
    sdfs_comm_state *cs = ctx->comm;
    if (cs->peers_mapped)
        for (int r = 0; r < ctx->nranks; ++r)
            if (r != ctx->rank && cs->peer_arena[r]) cudaIpcCloseMemHandle(cs->peer_arena[r]);
    // peers may still be storing into this arena: a collective barrier first (skipped when the exchange state is
    // poisoned by a peer timeout - the peer is gone - or when SDFS_COMM_DESTROY_BARRIER=0)
    static const bool barrier_on = !(getenv("SDFS_COMM_DESTROY_BARRIER") && atoi(getenv("SDFS_COMM_DESTROY_BARRIER")) == 0);
    if (cs->comm && cs->peers_mapped && barrier_on && ctx->h_status && !*(volatile long long *)ctx_h_abort(ctx)) {
        double *tmp = (double *)ctx->d_status + 256;
        if (g_nccl.AllReduce(tmp, tmp, 1, ncclFloat64, ncclSumOp, cs->comm, ctx->stream) == 0) cudaStreamSynchronize(ctx->stream);
    }
    if (cs->arena) cudaFree(cs->arena);
    if (cs->comm) g_nccl.CommDestroy(cs->comm);
    delete cs;
    ctx->comm = nullptr;
    return SDFS_OK;
}

// Row partition shared by every component: rank g owns [g*chunk, min(N,(g+1)*chunk)).
static inline void rank_rows(int64_t N, int nranks, int rank, int64_t *rb, int64_t *re) {
    const int64_t chunk = (N + nranks - 1) / nranks;
    int64_t b = chunk * rank, e = b + chunk;
    if (b > N) b = N;
    if (e > N) e = N;
    *rb = b; *re = e;
}

// In-place all-gather of a full-length vector: rank r owns rows [rb[r], re[r]) (contiguous, in rank order).
int comm_allgather_parts(sdfs_ctx *ctx, double *d_vec, const int64_t *rb, const int64_t *re) {
    if (ctx->nranks <= 1) return SDFS_OK;
    if (!ctx->comm || !ctx->comm->comm) return sdfs_set_error(ctx, SDFS_ERR_COMM, "communicator not initialised");
    bool equal = rb[0] == 0;
    for (int r = 1; r < ctx->nranks; ++r) equal = equal && (re[r] - rb[r] == re[0] - rb[0]) && rb[r] == re[r - 1];
    if (equal) {
        NCCL_TRY(ctx, g_nccl.AllGather(d_vec + rb[ctx->rank], d_vec, (size_t)(re[0] - rb[0]), ncclFloat64, ctx->comm->comm, ctx->stream));
    } else {
        NCCL_TRY(ctx, g_nccl.GroupStart());
        for (int r = 0; r < ctx->nranks; ++r)
            if (re[r] > rb[r])
                NCCL_TRY(ctx, g_nccl.Broadcast(d_vec + rb[r], d_vec + rb[r], (size_t)(re[r] - rb[r]), ncclFloat64, r, ctx->comm->comm, ctx->stream));
        NCCL_TRY(ctx, g_nccl.GroupEnd());
    }
    return SDFS_OK;
}

// ... with the default row partition of a vector of length N
int comm_allgather_rows(sdfs_ctx *ctx, double *d_vec, int64_t N) {
    int64_t rb[SDFS_MAX_RANKS], re[SDFS_MAX_RANKS];
    for (int r = 0; r < ctx->nranks; ++r) rank_rows(N, ctx->nranks, r, &rb[r], &re[r]);
    return comm_allgather_parts(ctx, d_vec, rb, re);
}


extern "C" {

int sdfs_comm_unique_id(void *h_id128) {
    if (!h_id128) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_comm_unique_id: NULL");
    int rc = nccl_load(nullptr);
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(h_id128, &id, 128);
    return SDFS_OK;
}

int sdfs_comm_init(sdfs_ctx *ctx, int rank, int nranks, const void *h_id128) {
    ARG_CHECK(ctx, ctx && h_id128 && nranks >= 1 && nranks <= SDFS_MAX_RANKS && rank >= 0 && rank < nranks);
    ARG_CHECK(ctx, ctx->comm == nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = nccl_load(ctx);
    if (rc) return rc;
    sdfs_comm_state *cs = new sdfs_comm_state();
    ncclUniqueId id;
    memcpy(&id, h_id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&cs->comm, nranks, id, rank);
    if (r != 0) {
        delete cs;
        return sdfs_set_error(ctx, SDFS_ERR_COMM, "ncclCommInitRank(rank %d/%d): %s", rank, nranks, g_nccl.GetErrorString(r));
    }
    ctx->comm = cs;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return SDFS_OK;
}

int sdfs_comm_rank(sdfs_ctx *ctx, int *rank, int *nranks) {
    ARG_CHECK(ctx, ctx != nullptr);
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return SDFS_OK;
}

int sdfs_comm_allgather_f64(sdfs_ctx *ctx, double *d_buf, int64_t count_per_rank) {
    ARG_CHECK(ctx, ctx && d_buf && count_per_rank >= 0);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->nranks <= 1) return SDFS_OK;
    return comm_allgather_rows(ctx, d_buf, count_per_rank * ctx->nranks);
}

int sdfs_comm_barrier(sdfs_ctx *ctx) {
    ARG_CHECK(ctx, ctx != nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (ctx->nranks > 1) {
        if (!ctx->comm) return sdfs_set_error(ctx, SDFS_ERR_COMM, "communicator not initialised");
        double *tmp = (double *)ctx->d_status + 256;   // scratch inside the status page
        NCCL_TRY(ctx, g_nccl.AllReduce(tmp, tmp, 1, ncclFloat64, ncclSumOp, ctx->comm->comm, ctx->stream));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFS_OK;
}

int sdfs_comm_arena_export(sdfs_ctx *ctx, int64_t max_N, void *h_handle64) {
    ARG_CHECK(ctx, ctx && h_handle64 && max_N >= 1);
    if (!ctx->comm) return sdfs_set_error(ctx, SDFS_ERR_COMM, "sdfs_comm_init first");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    sdfs_comm_state *cs = ctx->comm;
    ARG_CHECK(ctx, cs->arena == nullptr);
    cs->arena_bytes = arena_bytes_for(max_N);
    cs->arena_maxN = max_N;
    CUDA_TRY(ctx, cudaMalloc(&cs->arena, cs->arena_bytes));
    CUDA_TRY(ctx, cudaMemset(cs->arena, 0, cs->arena_bytes));
    CUDA_TRY(ctx, cudaDeviceSynchronize());   // zeroed before any peer can see the handle
    cudaIpcMemHandle_t h;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&h, cs->arena));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(h_handle64, &h, 64);
    return SDFS_OK;
}

int sdfs_comm_arena_import(sdfs_ctx *ctx, const void *h_handles64_all) {
    ARG_CHECK(ctx, ctx && h_handles64_all);
    if (!ctx->comm || !ctx->comm->arena) return sdfs_set_error(ctx, SDFS_ERR_COMM, "sdfs_comm_arena_export first");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    sdfs_comm_state *cs = ctx->comm;
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) { cs->peer_arena[r] = cs->arena; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)h_handles64_all + 64 * r, 64);
        CUDA_TRY(ctx, cudaIpcOpenMemHandle(&cs->peer_arena[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    cs->peers_mapped = true;
    return SDFS_OK;
}

}  // extern "C"

// accessors used by loops.cu
bool comm_peers_ready(sdfs_ctx *ctx) { return ctx->comm && ctx->comm->peers_mapped; }
void *comm_peer_arena(sdfs_ctx *ctx, int r) { return ctx->comm ? ctx->comm->peer_arena[r] : nullptr; }
int64_t comm_arena_maxN(sdfs_ctx *ctx) { return ctx->comm ? ctx->comm->arena_maxN : 0; }
unsigned long long *comm_epoch(sdfs_ctx *ctx) { return ctx->comm ? &ctx->comm->epoch : nullptr; }
