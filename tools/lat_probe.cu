// Latency micro-probe (single warp unless noted): DFMA chain, exp/log/pow, 64-bit shuffle reduce, barrier.
#include <cuda_runtime.h>
#include <stdio.h>
#include <math.h>
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void probe(double *out, long long *cyc, double seed, double theta) {
    const int tid = threadIdx.x;
    double x = seed + tid * 1e-3;
    long long t0, t1;
    // 1. dependent DFMA chain (256 long)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 32; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x = fma(x, 1.0000001, 1e-9);
    }
    t1 = clock64();
    if (tid == 0) cyc[0] = (t1 - t0) / 256;
    // 2. exp
    double y = x * 1e-3;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) y = exp(y * 1e-3) - 0.99;
    t1 = clock64();
    if (tid == 0) cyc[1] = (t1 - t0) / 64;
    // 3. log
    double z = 2.0 + y;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) z = log(z + 700.0);
    t1 = clock64();
    if (tid == 0) cyc[2] = (t1 - t0) / 64;
    // 4. pow
    double p = 700.0 + z;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) p = pow(p, theta) * 1e47 + 700.0;
    t1 = clock64();
    if (tid == 0) cyc[3] = (t1 - t0) / 64;
    // 5. warp_sum (5 x 64-bit shuffle + add)
    double s = p;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) s = warp_sum(s) * (1.0 / 32.0);
    t1 = clock64();
    if (tid == 0) cyc[4] = (t1 - t0) / 64;
    // 6. __syncthreads (whole block)
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64();
    if (tid == 0) cyc[5] = (t1 - t0) / 64;
    // 7. shared memory load-use latency (pointer chase)
    __shared__ int chain[64];
    if (tid < 64) chain[tid] = (tid + 1) & 63;
    __syncthreads();
    int q = tid & 63;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) q = chain[q];
    t1 = clock64();
    if (tid == 0) cyc[6] = (t1 - t0) / 64;
    // 8. exp(theta*log(x)) pair
    double e = 700.0 + q;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) e = exp(theta * log(e)) * 1e47 + 700.0;
    t1 = clock64();
    if (tid == 0) cyc[7] = (t1 - t0) / 64;
    out[tid] = x + y + z + p + s + q + e;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
    const char *names[8] = {"DFMA dependent", "exp", "log", "pow", "warp_sum(f64)", "__syncthreads", "LDS chase", "exp(t*log)"};
    for (int threads : {32, 512, 1024}) {
        probe<<<1, threads>>>(out, cyc, 1.0, -16.02);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads=%4d:", threads);
        for (int i = 0; i < 8; ++i) printf("  %s=%lld", names[i], h[i]);
        printf("\n");
    }
    return 0;
}
