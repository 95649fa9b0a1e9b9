// What feeds DMMA m8n8k4 at full rate on B200?  (register-only peak: tools/pipe_probe.cu, 36.97 TFLOP/s)
//   mode 0: 7 accumulator chains, A and B operands from DISTINCT registers (no operand reuse), no memory
//   mode 1: as 0 but every B operand comes from its own LDS.64 (the factor-form contraction: 1 LDS per DMMA)
//   mode 2: B operands loaded once into registers and reused across 4 independent fibre tiles (1 LDS per 4 DMMA)
//   mode 3: as 1 but A from LDS too (A once per 7 DMMAs), i.e. the shipped inner loop
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#define IT 7
#define KT 14
#define PITCH 60
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(int iters, double seed, double *out) {
    __shared__ double smat[56 * PITCH];
    __shared__ double sa[4][KT][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
    for (int e = threadIdx.x; e < 56 * PITCH; e += blockDim.x) smat[e] = seed + e * 1e-9;
    for (int kt = 0; kt < KT; ++kt) sa[warp & 3][kt][lane] = seed * kt + lane * 1e-9;
    __syncthreads();
    const double *brow = smat + g * PITCH + q;
    double acc = 0.0;
    if (MODE == 2) {
        // warp owns 2 output tiles (B in registers: 2 x KT doubles), 4 fibre tiles in flight
        double b[2][KT];
#pragma unroll
        for (int it = 0; it < 2; ++it)
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) b[it][kt] = brow[it * 8 * PITCH + kt * 4];
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            double c[4][2][2];
#pragma unroll
            for (int t = 0; t < 4; ++t) c[t][0][0] = c[t][0][1] = c[t][1][0] = c[t][1][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double a = sa[(warp + t) & 3][kt][lane];
                    dmma884(c[t][0][0], c[t][0][1], a, b[0][kt]);
                    dmma884(c[t][1][0], c[t][1][1], a, b[1][kt]);
                }
#pragma unroll
            for (int t = 0; t < 4; ++t) acc += c[t][0][0] + c[t][0][1] + c[t][1][0] + c[t][1][1];
        }
    } else if (MODE == 4) {
        // two fibre tiles per warp iteration: 14 independent accumulator chains, each B fragment feeds two DMMAs
#pragma unroll 1
        for (int i = 0; i < iters; i += 2) {
            double c[2][IT][2];
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int it = 0; it < IT; ++it) c[t][it][0] = c[t][it][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const double a0 = sa[warp & 3][kt][lane], a1 = sa[(warp + 1) & 3][kt][lane];
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    const double b = brow[it * 8 * PITCH + kt * 4];
                    dmma884(c[0][it][0], c[0][it][1], a0, b);
                    dmma884(c[1][it][0], c[1][it][1], a1, b);
                }
            }
#pragma unroll
            for (int it = 0; it < IT; ++it) acc += c[0][it][0] + c[0][it][1] + c[1][it][0] + c[1][it][1];
            if (acc == 1.2345) sa[warp & 3][0][lane] = acc;
        }
    } else {
        double areg[KT];
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) areg[kt] = sa[warp & 3][kt][lane];
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            double c[IT][2];
#pragma unroll
            for (int it = 0; it < IT; ++it) c[it][0] = c[it][1] = 0.0;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                const double a = (MODE == 3) ? sa[warp & 3][kt][lane] : areg[kt];
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    const double b = (MODE == 0) ? areg[(kt + it + 1) % KT] : brow[it * 8 * PITCH + kt * 4];
                    dmma884(c[it][0], c[it][1], a, b);
                }
            }
#pragma unroll
            for (int it = 0; it < IT; ++it) acc += c[it][0] + c[it][1];
            if (MODE == 3 && acc == 1.2345) sa[warp & 3][0][lane] = acc;      // keep the A loads inside the loop
        }
    }
    if (acc == 12345.678) out[0] = acc;
}
template <int MODE> static void run(const char *name, int grid, double *d_out, int threads = 256) {
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, threads>>>(iters / 8, 0.5, d_out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<grid, threads>>>(iters, 0.5, d_out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double dm = (double)grid * (threads / 32) * iters * (MODE == 2 ? 4 * 2 * KT : IT * KT);
    printf(" \"%s\": {\"ms\": %.3f, \"tflops\": %.2f},\n", name, best, dm * 512.0 / best * 1e-9);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *d_out; cudaMalloc(&d_out, 64);
    const int grid = p.multiProcessorCount * 2;
    printf("{\n");
    run<0>("distinct_registers_no_memory", grid, d_out);
    run<1>("B_from_LDS_per_DMMA", grid, d_out);
    run<3>("B_from_LDS_per_DMMA_A_from_LDS_per_7", grid, d_out);
    run<2>("B_in_registers_4_fibre_tiles_A_from_LDS", grid, d_out);
    run<4>("two_fibre_tiles_per_iteration_14_chains_16_warps_per_sm", grid, d_out);
    // how many warps does it take to fill the pipe?  (one CTA per SM, 4 / 8 warps = 1 / 2 per scheduler)
    run<3>("shipped_loop_1_warp_per_scheduler", grid / 2, d_out, 128);
    run<4>("two_tiles_1_warp_per_scheduler", grid / 2, d_out, 128);
    run<3>("shipped_loop_2_warps_per_scheduler", grid / 2, d_out, 256);
    run<4>("two_tiles_2_warps_per_scheduler", grid / 2, d_out, 256);
    printf(" \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
