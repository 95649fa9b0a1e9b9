// Context, memory, error handling and DLPack entry points of libsdfs_b200.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

thread_local std::string g_last_error;

int sdfs_set_error(sdfs_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_last_error = buf;
    (void)cudaGetLastError();   // a handled failure (e.g. out of memory) must not leak into the next call
    return code;
}

int comm_destroy(sdfs_ctx *ctx);   // comm.cu

extern "C" {

int sdfs_abi_version(void) { return SDFS_ABI_VERSION; }
const char *sdfs_version_string(void) { return "sdfs_b200 0.1 (sm_100a, fp64)"; }

int sdfs_ctx_create(int device, sdfs_ctx **out) {
    if (!out) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return sdfs_set_error(nullptr, SDFS_ERR_CUDA,
                              "sdfs_ctx_create: no CUDA device available (%s); this library has no "
                              "CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev)
        return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_ctx_create: device %d out of range [0,%d)",
                              device, ndev);
    sdfs_ctx *ctx = new sdfs_ctx();
    ctx->device = device;
    // every failure below releases what was created so far (sdfs_ctx_destroy tolerates null members)
    auto init = [&]() -> int {
        CUDA_TRY(nullptr, cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return sdfs_set_error(nullptr, SDFS_ERR_UNSUPPORTED,
                                  "sdfs_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                                  device, prop.major, prop.minor);
        ctx->sm_count = prop.multiProcessorCount;
        ctx->coop_supported = prop.cooperativeLaunch;
        CUDA_TRY(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        {
            cudaMemPool_t pool;
            CUDA_TRY(nullptr, cudaDeviceGetDefaultMemPool(&pool, device));
            uint64_t keep = ~0ull;   // keep freed blocks cached in the pool
            CUDA_TRY(nullptr, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        CUDA_TRY(nullptr, cudaEventCreate(&ctx->ev0));
        CUDA_TRY(nullptr, cudaEventCreate(&ctx->ev1));
        CUDA_TRY(nullptr, cudaMalloc(&ctx->d_status, 4096));
        CUDA_TRY(nullptr, cudaMemset(ctx->d_status, 0, 4096));
        CUDA_TRY(nullptr, cudaMallocHost(&ctx->h_status, 4096));
        memset(ctx->h_status, 0, 4096);
        return SDFS_OK;
    };
    const int rc = init();
    if (rc != SDFS_OK) {
        const std::string msg = g_last_error;        // sdfs_ctx_destroy must not clobber the reason
        sdfs_ctx_destroy(ctx);
        g_last_error = msg;
        return rc;
    }
    *out = ctx;
    return SDFS_OK;
}

int sdfs_ctx_destroy(sdfs_ctx *ctx) {
    if (!ctx) return SDFS_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_destroy(ctx);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t ev : ctx->prof_ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SDFS_OK;
}

const char *sdfs_last_error(sdfs_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int sdfs_ctx_sync(sdfs_ctx *ctx) {
    ARG_CHECK(ctx, ctx != nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(volatile long long *)ctx_h_abort(ctx))
        return sdfs_set_error(ctx, SDFS_ERR_TIMEOUT, "a peer rank did not reach a fused exchange (30 s device-side timeout); the context's exchange state is poisoned: destroy and re-create the contexts of all ranks");
    return SDFS_OK;
}

int sdfs_ctx_device_sync(sdfs_ctx *ctx) {
    ARG_CHECK(ctx, ctx != nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    return SDFS_OK;
}

int sdfs_ctx_device(sdfs_ctx *ctx, int *device, int *sm_count, size_t *free_bytes, size_t *total_bytes) {
    ARG_CHECK(ctx, ctx != nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (device) *device = ctx->device;
    if (sm_count) *sm_count = ctx->sm_count;
    size_t f = 0, t = 0;
    CUDA_TRY(ctx, cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return SDFS_OK;
}

int64_t sdfs_ctx_launch_count(sdfs_ctx *ctx) { return ctx ? ctx->launches : -1; }

int sdfs_timer_start(sdfs_ctx *ctx) {
    ARG_CHECK(ctx, ctx != nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return SDFS_OK;
}

int sdfs_timer_stop_ms(sdfs_ctx *ctx, double *ms) {
    ARG_CHECK(ctx, ctx != nullptr && ms != nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    if (*(volatile long long *)ctx_h_abort(ctx))
        return sdfs_set_error(ctx, SDFS_ERR_TIMEOUT, "a peer rank did not reach a fused exchange (30 s device-side timeout); the context's exchange state is poisoned: destroy and re-create the contexts of all ranks");
    float f = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    *ms = (double)f;
    return SDFS_OK;
}

int sdfs_prof_enable(sdfs_ctx *ctx, int max_launches) {
    ARG_CHECK(ctx, ctx != nullptr && max_launches >= 0 && max_launches <= (1 << 20));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    while (ctx->prof_ev.size() < (size_t)max_launches * 2) {
        cudaEvent_t ev;
        CUDA_TRY(ctx, cudaEventCreate(&ev));
        ctx->prof_ev.push_back(ev);
    }
    ctx->prof_used = 0;
    ctx->prof_on = max_launches > 0;
    return SDFS_OK;
}

int sdfs_prof_read(sdfs_ctx *ctx, double *total_ms, int64_t *launches) {
    ARG_CHECK(ctx, ctx != nullptr && total_ms && launches);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[i], ctx->prof_ev[i + 1]));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)(ctx->prof_used / 2);
    ctx->prof_used = 0;
    return SDFS_OK;
}

int sdfs_malloc(sdfs_ctx *ctx, size_t bytes, void **d_ptr) {
    ARG_CHECK(ctx, ctx != nullptr && d_ptr != nullptr);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    *d_ptr = nullptr;
    if (bytes == 0) bytes = 8;
    // stream-ordered pool: no device-wide synchronisation per allocation (cudaMalloc/cudaFree
    // cost ~100 ms each once an 88 GB operator is mapped)
    CUDA_TRY(ctx, cudaMallocAsync(d_ptr, bytes, ctx->stream));
    return SDFS_OK;
}

int sdfs_free(sdfs_ctx *ctx, void *d_ptr) {
    ARG_CHECK(ctx, ctx != nullptr);
    if (!d_ptr) return SDFS_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaFreeAsync(d_ptr, ctx->stream));
    return SDFS_OK;
}

int sdfs_memset(sdfs_ctx *ctx, void *d_ptr, int value, size_t bytes) {
    ARG_CHECK(ctx, ctx != nullptr && (d_ptr != nullptr || bytes == 0));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemsetAsync(d_ptr, value, bytes, ctx->stream));
    return SDFS_OK;
}

int sdfs_h2d(sdfs_ctx *ctx, void *d_dst, const void *h_src, size_t bytes) {
    ARG_CHECK(ctx, ctx != nullptr && (bytes == 0 || (d_dst && h_src)));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return SDFS_OK;
}

int sdfs_d2h(sdfs_ctx *ctx, void *h_dst, const void *d_src, size_t bytes) {
    ARG_CHECK(ctx, ctx != nullptr && (bytes == 0 || (h_dst && d_src)));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(volatile long long *)ctx_h_abort(ctx))
        return sdfs_set_error(ctx, SDFS_ERR_TIMEOUT, "a peer rank did not reach a fused exchange (30 s device-side timeout); the context's exchange state is poisoned: destroy and re-create the contexts of all ranks");
    return SDFS_OK;
}

int sdfs_d2d(sdfs_ctx *ctx, void *d_dst, const void *d_src, size_t bytes) {
    ARG_CHECK(ctx, ctx != nullptr && (bytes == 0 || (d_dst && d_src)));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return SDFS_OK;
}

__global__ void k_fill(double *p, double v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

int sdfs_fill_f64(sdfs_ctx *ctx, double *d_ptr, double value, int64_t n) {
    ARG_CHECK(ctx, ctx != nullptr && (d_ptr != nullptr || n == 0) && n >= 0);
    if (n == 0) return SDFS_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int64_t g = (n + 255) / 256;
    if (g > 4096) g = 4096;
    k_fill<<<(int)g, 256, 0, ctx->stream>>>(d_ptr, value, n);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

int sdfs_host_alloc_pinned(size_t bytes, void **h_ptr) {
    if (!h_ptr) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_host_alloc_pinned: NULL out");
    cudaError_t e = cudaMallocHost(h_ptr, bytes ? bytes : 8);
    if (e != cudaSuccess)
        return sdfs_set_error(nullptr, SDFS_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return SDFS_OK;
}

int sdfs_host_free_pinned(void *h_ptr) {
    if (h_ptr) cudaFreeHost(h_ptr);
    return SDFS_OK;
}

// ---------------------------------------------------------------------------
// DLPack (ABI of dlpack.h v0.8: DLManagedTensor)
// ---------------------------------------------------------------------------
typedef struct { int32_t device_type; int32_t device_id; } DLDevice_;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType_;
typedef struct {
    void *data; DLDevice_ device; int32_t ndim; DLDataType_ dtype;
    int64_t *shape; int64_t *strides; uint64_t byte_offset;
} DLTensor_;
typedef struct DLManagedTensor_ {
    DLTensor_ dl_tensor; void *manager_ctx; void (*deleter)(struct DLManagedTensor_ *self);
} DLManagedTensor_;

struct ExportCtx { int64_t shape[8]; void *token; void (*release)(void *); };

static void export_deleter(DLManagedTensor_ *self) {
    if (!self) return;
    ExportCtx *ec = (ExportCtx *)self->manager_ctx;
    if (ec) {
        if (ec->release) ec->release(ec->token);
        free(ec);
    }
    free(self);
}

int sdfs_dlpack_export(sdfs_ctx *ctx, void *d_ptr, int ndim, const int64_t *shape, void *owner_token,
                       void (*release)(void *), void **dl_managed_tensor) {
    ARG_CHECK(ctx, ctx && d_ptr && shape && dl_managed_tensor && ndim >= 0 && ndim <= 8);
    DLManagedTensor_ *mt = (DLManagedTensor_ *)calloc(1, sizeof(DLManagedTensor_));
    ExportCtx *ec = (ExportCtx *)calloc(1, sizeof(ExportCtx));
    if (!mt || !ec) return sdfs_set_error(ctx, SDFS_ERR_NOMEM, "dlpack export: host allocation failed");
    for (int i = 0; i < ndim; ++i) ec->shape[i] = shape[i];
    ec->token = owner_token;
    ec->release = release;
    mt->dl_tensor.data = d_ptr;
    mt->dl_tensor.device.device_type = 2;  // kDLCUDA
    mt->dl_tensor.device.device_id = ctx->device;
    mt->dl_tensor.ndim = ndim;
    mt->dl_tensor.dtype.code = 2;          // kDLFloat
    mt->dl_tensor.dtype.bits = 64;
    mt->dl_tensor.dtype.lanes = 1;
    mt->dl_tensor.shape = ec->shape;
    mt->dl_tensor.strides = nullptr;       // C-contiguous
    mt->dl_tensor.byte_offset = 0;
    mt->manager_ctx = ec;
    mt->deleter = export_deleter;
    *dl_managed_tensor = mt;
    return SDFS_OK;
}

int sdfs_dlpack_import(void *dl_managed_tensor, void **d_ptr, int *ndim, int64_t *shape8, int *device,
                       int64_t *n_elems) {
    if (!dl_managed_tensor || !d_ptr || !ndim || !shape8 || !device || !n_elems)
        return sdfs_set_error(nullptr, SDFS_ERR_ARG, "dlpack import: NULL argument");
    DLManagedTensor_ *mt = (DLManagedTensor_ *)dl_managed_tensor;
    const DLTensor_ &t = mt->dl_tensor;
    if (t.device.device_type != 2 && t.device.device_type != 13 /* kDLCUDAManaged */)
        return sdfs_set_error(nullptr, SDFS_ERR_UNSUPPORTED,
                              "dlpack import: device_type %d is not CUDA (no CPU path exists)",
                              t.device.device_type);
    if (t.dtype.code != 2 || t.dtype.bits != 64 || t.dtype.lanes != 1)
        return sdfs_set_error(nullptr, SDFS_ERR_UNSUPPORTED, "dlpack import: dtype must be float64");
    if (t.ndim < 0 || t.ndim > 8)
        return sdfs_set_error(nullptr, SDFS_ERR_UNSUPPORTED, "dlpack import: ndim %d > 8", t.ndim);
    int64_t n = 1;
    for (int i = 0; i < t.ndim; ++i) n *= t.shape[i];
    if (t.strides) {
        int64_t expect = 1;
        for (int i = t.ndim - 1; i >= 0; --i) {
            if (t.shape[i] != 1 && t.strides[i] != expect)
                return sdfs_set_error(nullptr, SDFS_ERR_UNSUPPORTED, "dlpack import: tensor is not C-contiguous");
            expect *= t.shape[i];
        }
    }
    *d_ptr = (char *)t.data + t.byte_offset;
    *ndim = t.ndim;
    for (int i = 0; i < t.ndim; ++i) shape8[i] = t.shape[i];
    *device = t.device.device_id;
    *n_elems = n;
    return SDFS_OK;
}

void sdfs_dlpack_call_deleter(void *dl_managed_tensor) {
    DLManagedTensor_ *mt = (DLManagedTensor_ *)dl_managed_tensor;
    if (mt && mt->deleter) mt->deleter(mt);
}

}  // extern "C"
