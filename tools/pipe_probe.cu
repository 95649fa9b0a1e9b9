// Throughput micro-probe for the fp64 pipes of one B200 (register-only, no memory traffic):
//   1. DMMA m8n8k4 peak (the denominator of every "tensor pipe" figure in profiles/)
//   2. DFMA peak
//   3. DMMA and DFMA issued by different warps of the same SM at the same time: do the rates add?
//   4. DMMA next to the exp(e log x) power form (the fused prologue/epilogue mix)
//   5. cost of a cooperative grid barrier at the loop kernels' launch shape (2 x 148 CTAs x 288 threads)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode bit 0: even warps run DMMA; bit 1: odd warps run DFMA; bit 2: odd warps run exp(e log x)
// (when only one kind is selected, every warp runs it)
__global__ void __launch_bounds__(256) k_mix(int mode, int iters, int iters_b, double seed, double *out) {
    const int warp = threadIdx.x >> 5;
    const bool only_dmma = mode == 1, only_dfma = mode == 2, only_pow = mode == 4;
    const bool do_dmma = only_dmma || (!only_dfma && !only_pow && (mode & 1) && (warp & 1) == 0);
    const bool do_dfma = only_dfma || (!only_dmma && !only_pow && (mode & 2) && (warp & 1) == 1);
    const bool do_pow = only_pow || (!only_dmma && !only_dfma && (mode & 4) && (warp & 1) == 1);
    double acc = 0.0;
    if (do_dmma) {
        double c[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
        const double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);      // 8 independent chains
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += c[i][0] + c[i][1];
    } else if (do_dfma) {
        double c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = seed * i;
        const double a = 1.0 - 1e-9 * seed, b = 1e-12;
#pragma unroll 1
        for (int it = 0; it < iters_b; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);               // 256 DFMA per lane = 32 DMMA worth
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc += c[i];
    } else if (do_pow) {
        double x[4] = {700.0 + seed, 710.0 + seed, 720.0 + seed, 730.0 + seed};
#pragma unroll 1
        for (int it = 0; it < iters_b; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = 700.0 + exp(-16.02 * log(x[i])) * 1e40;
        }
        acc = x[0] + x[1] + x[2] + x[3];
    }
    if (acc == 12345.678) out[0] = acc;
}

__global__ void __launch_bounds__(288, 2) k_gsync(int reps, long long *cyc) {
    cg::grid_group grid = cg::this_grid();
    grid.sync();
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) grid.sync();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = (t1 - t0) / reps;
}

static float time_mix(int mode, int iters, int iters_b, int grid, double *d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<<<grid, 256>>>(mode, iters / 8, iters_b / 8, 0.5, d_out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        k_mix<<<grid, 256>>>(mode, iters, iters_b, 0.5, d_out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *d_out; cudaMalloc(&d_out, 64);
    long long *d_cyc; cudaMalloc(&d_cyc, 64);
    const int iters = 20000, grid = sms * 2;              // 16 warps per SM
    const double warps = (double)grid * 8;
    const int iters_p = iters / 16;
    const float t_dmma = time_mix(1, iters, 0, grid, d_out);
    const float t_dfma = time_mix(2, 0, iters, grid, d_out);
    const float t_pow = time_mix(4, 0, iters_p, grid, d_out);
    // half the warps each, every warp with the same work as in its solo run: a shared pipe gives
    // (t_a + t_b) / 2, separate pipes give max(t_a, t_b) / 2
    const float t_both = time_mix(3, iters, iters, grid, d_out);
    const int iters_pm = (int)(iters_p * (double)t_dmma / t_pow);     // pow work scaled to the DMMA duration
    const float t_pow_m = time_mix(4, 0, iters_pm, grid, d_out);
    const float t_dp = time_mix(5, iters, iters_pm, grid, d_out);
    const double fl_dmma = warps * iters * 32.0 * 512.0;    // 32 DMMA x (8x8x4 x 2 flop)
    const double fl_dfma = warps * iters * 256.0 * 32 * 2.0;
    printf("{\"sms\": %d, \"dmma_only_ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_only_ms\": %.3f, \"dfma_tflops\": %.2f,\n",
           sms, t_dmma, fl_dmma / t_dmma * 1e-9, t_dfma, fl_dfma / t_dfma * 1e-9);
    printf(" \"half_dmma_half_dfma_ms\": %.3f, \"shared_pipe_would_be_ms\": %.3f, \"separate_pipes_would_be_ms\": %.3f,\n", t_both,
           (t_dmma + t_dfma) / 2, (t_dmma > t_dfma ? t_dmma : t_dfma) / 2);
    printf(" \"pow_only_ms\": %.3f, \"pow_per_s\": %.3e, \"pow_scaled_ms\": %.3f, \"half_dmma_half_pow_ms\": %.3f,\n", t_pow,
           warps * 32 * 4.0 * iters_p / (t_pow * 1e-3), t_pow_m, t_dp);
    // grid barrier
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gsync, 288, 0);
    for (int cps = 1; cps <= (per_sm < 2 ? per_sm : 2); ++cps) {
        int reps = 2000, g = sms * cps;
        void *args[] = {&reps, &d_cyc};
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaLaunchCooperativeKernel((void *)k_gsync, dim3(g), dim3(288), args, 0, 0);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        cudaLaunchCooperativeKernel((void *)k_gsync, dim3(g), dim3(288), args, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf(" \"grid_sync_us_%dcta_per_sm\": %.3f, \"grid_sync_cycles_%d\": %lld,\n", cps, ms * 1e3 / reps, cps, cyc);
    }
    // empty-kernel launch cadence (back-to-back launches on one stream)
    {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < 1000; ++i) k_mix<<<grid, 256>>>(0, 0, 0, 0.5, d_out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf(" \"empty_launch_us\": %.3f, \"err\": \"%s\"}\n", ms, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
