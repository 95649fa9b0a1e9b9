// What do the Krylov vector phases of the solver loops get out of HBM?  z = x + a y over 9.8 M-element vectors
// (3 streams: 2 reads + 1 write, 236 MB), persistent grid as in the loop kernels (2 CTAs x 256 threads per SM) vs a full
// occupancy launch, 8-byte vs 16-byte accesses, 1 / 4 / 8 independent elements per trip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_probe tools/stream_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int VU>
__global__ void __launch_bounds__(256) k8(long long n, double a, const double *x, const double *y, double *z) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long n0 = tid; n0 < n; n0 += VU * nth) {
        double xv[VU], yv[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) { const long long i = n0 + u * nth; if (i < n) { xv[u] = x[i]; yv[u] = y[i]; } }
#pragma unroll
        for (int u = 0; u < VU; ++u) { const long long i = n0 + u * nth; if (i < n) z[i] = xv[u] + a * yv[u]; }
    }
}
template <int VU>
__global__ void __launch_bounds__(256) k16(long long n2, double a, const double2 *x, const double2 *y, double2 *z) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long n0 = tid; n0 < n2; n0 += VU * nth) {
        double2 xv[VU], yv[VU];
#pragma unroll
        for (int u = 0; u < VU; ++u) { const long long i = n0 + u * nth; if (i < n2) { xv[u] = x[i]; yv[u] = y[i]; } }
#pragma unroll
        for (int u = 0; u < VU; ++u) { const long long i = n0 + u * nth; if (i < n2) z[i] = make_double2(xv[u].x + a * yv[u].x, xv[u].y + a * yv[u].y); }
    }
}
template <class F> static void timeit(const char *name, double bytes, F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf(" \"%s\": {\"us\": %.1f, \"GBps\": %.0f},\n", name, best * 1e3, bytes / best * 1e-6);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const long long n = 9834496;
    double *x, *y, *z; cudaMalloc(&x, n * 8); cudaMalloc(&y, n * 8); cudaMalloc(&z, n * 8);
    cudaMemset(x, 0, n * 8); cudaMemset(y, 0, n * 8);
    const double bytes = 3.0 * n * 8;
    const int pg = p.multiProcessorCount * 2, fg = (int)((n + 255) / 256);
    printf("{\n");
    timeit("persistent_8B_vu1", bytes, [&] { k8<1><<<pg, 256>>>(n, 0.5, x, y, z); });
    timeit("persistent_8B_vu4", bytes, [&] { k8<4><<<pg, 256>>>(n, 0.5, x, y, z); });
    timeit("persistent_8B_vu8", bytes, [&] { k8<8><<<pg, 256>>>(n, 0.5, x, y, z); });
    timeit("persistent_16B_vu2", bytes, [&] { k16<2><<<pg, 256>>>(n / 2, 0.5, (double2 *)x, (double2 *)y, (double2 *)z); });
    timeit("persistent_16B_vu4", bytes, [&] { k16<4><<<pg, 256>>>(n / 2, 0.5, (double2 *)x, (double2 *)y, (double2 *)z); });
    timeit("persistent_16B_vu8", bytes, [&] { k16<8><<<pg, 256>>>(n / 2, 0.5, (double2 *)x, (double2 *)y, (double2 *)z); });
    timeit("full_grid_8B_vu1", bytes, [&] { k8<1><<<fg, 256>>>(n, 0.5, x, y, z); });
    timeit("full_grid_16B_vu1", bytes, [&] { k16<1><<<fg / 2, 256>>>(n / 2, 0.5, (double2 *)x, (double2 *)y, (double2 *)z); });
    timeit("persistent4x_8B_vu4", bytes, [&] { k8<4><<<pg * 4, 256>>>(n, 0.5, x, y, z); });
    timeit("persistent4x_16B_vu4", bytes, [&] { k16<4><<<pg * 4, 256>>>(n / 2, 0.5, (double2 *)x, (double2 *)y, (double2 *)z); });
    printf(" \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
