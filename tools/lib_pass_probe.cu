// A/B harness: the library's own dense pass (csrc/rowdot.cuh) run in a bare CUDA program,
// back to back, on synthetic data -- isolates kernel code from the Python/ctypes environment.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/lib_pass_probe tools/lib_pass_probe.cu
#include "../sdfs_via_autodiff_b200/csrc/rowdot.cuh"
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void k_fillP(double *P, int64_t N, int64_t ld) {
    const int64_t total = N * ld;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e / ld, j = e % ld;
        P[e] = (j < N) ? (double)((i * 31 + j * 17) % 97) / 97.0 : 0.0;
    }
}
__global__ void k_fillx(double *x, int64_t N, int64_t ld) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < ld; j += (int64_t)gridDim.x * blockDim.x)
        x[j] = (j < N) ? 1.0 + (double)(j % 13) : 0.0;
}
template <int VARIANT>
__global__ void __launch_bounds__(SDFS_THREADS, 1) k_lib_pass(const __grid_constant__ DenseView dv, const double *x, double *y) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    RowPipe<1> *rp = reinterpret_cast<RowPipe<1> *>(dyn_smem);
    PipeState st;
    pipe_init(rp, st);
    dense_pass_tma<1>(dv, x, x, rp, st, [&](int64_t n, double s0, double) { y[n] = s0; });
}
int main(int argc, char **argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 38416;
    const double secs = argc > 2 ? atof(argv[2]) : 1.5;
    const int64_t ld = (N + 63) / 64 * 64;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double *P, *x, *y;
    if (cudaMalloc(&P, (size_t)N * ld * 8) != cudaSuccess) { printf("alloc failed\n"); return 2; }
    CK(cudaMalloc(&x, (ld + 512) * 8)); CK(cudaMalloc(&y, N * 8));
    k_fillP<<<sms * 16, 256>>>(P, N, ld);
    k_fillx<<<64, 256>>>(x, N, ld + 512);
    CK(cudaDeviceSynchronize());
    DenseView dv{};
    dv.P = P; dv.N = N; dv.ld = ld; dv.row_begin = 0; dv.row_end = N; dv.vec2 = 1;
    {
        typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                          const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)N}; cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
        cuuint32_t box[2] = {TCW, TR}; cuuint32_t es[2] = {1, 1};
        if (((EncodeTiledFn)fn)(&dv.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, P, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            printf("tensor map failed\n"); return 3;
        }
    }
    CK(cudaFuncSetAttribute(k_lib_pass<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RowPipe<1>)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = ((double)N * N * 8 + 2.0 * N * 8) / 1e9;
    for (int round = 0; round < 2; ++round) {
        k_lib_pass<0><<<sms, SDFS_THREADS, sizeof(RowPipe<1>)>>>(dv, x, y);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k_lib_pass<0><<<sms, SDFS_THREADS, sizeof(RowPipe<1>)>>>(dv, x, y); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms1; cudaEventElapsedTime(&ms1, e0, e1);
        const int n = (int)(secs * 1e3 / ms1) + 1;
        cudaEventRecord(e0);
        for (int i = 0; i < n; ++i) k_lib_pass<0><<<sms, SDFS_THREADS, sizeof(RowPipe<1>)>>>(dv, x, y);
        cudaEventRecord(e1);
        char buf[256] = {0};
        FILE *f = popen("sleep 0.6; nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits | head -1", "r");
        if (f) { if (!fgets(buf, sizeof(buf), f)) buf[0] = 0; pclose(f); }
        for (char *c = buf; *c; ++c) if (*c == '\n') *c = 0;
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= n;
        printf("LIB dense_pass_tma bare: first %.3f ms | avg %.3f ms over %d launches %8.1f GB/s | sm_mhz,power: %s\n", ms1, ms, n, gb / ms * 1e3, buf);
    }
    double h[4]; CK(cudaMemcpy(h, y, 32, cudaMemcpyDeviceToHost));
    printf("y[0..3] = %.6f %.6f %.6f %.6f\n", h[0], h[1], h[2], h[3]);
    return 0;
}
