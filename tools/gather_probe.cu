// Does the 64-byte granularity of the factor-form contraction's gathers / scatters cap its memory rate?
// Copy of a 9.8 M-element vector with the access pattern of a middle mode at (56,)^4 (h_c: stride 3136 elements):
// a warp moves a "tile" of F contiguous fibres x 56 k steps; per instruction the 32 lanes cover 32/F' k rows of
// F contiguous elements.  F = 8 with 8-byte lanes (shipped: 64-byte pieces), F = 16 with 8-byte or 16-byte lanes
// (128-byte pieces), F = 32 (256-byte pieces), and the fully coalesced copy for reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
template <int F, int W>   // F fibres per tile, W doubles per lane access (1 or 2)
__global__ void __launch_bounds__(256, 2) k(const double *__restrict__ in, double *__restrict__ out, long long nfib, int n, long long stride,
                                            long long inner /* fibres contiguous before jumping by n*stride */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int LPR = F / W;            // lanes per k row
    constexpr int RPI = 32 / LPR;         // k rows per instruction
    const int lf = (lane % LPR) * W, lr = lane / LPR;
    const long long tiles = nfib / F;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + warp, nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long t = wid; t < tiles; t += nw) {
        const long long f0 = t * F;
        const long long base = (f0 / inner) * inner * n + (f0 % inner);     // fibre f: outer block * n * stride + inner offset (stride == inner)
        double v[64 / RPI][W];
#pragma unroll
        for (int r = 0; r < 64 / RPI; ++r) {
            const int krow = r * RPI + lr;
            if (krow < n) {
                const double *p = in + base + lf + (long long)krow * stride;
                if (W == 2) { const double2 x = *reinterpret_cast<const double2 *>(p); v[r][0] = x.x; v[r][W - 1] = x.y; }
                else v[r][0] = *p;
            }
        }
#pragma unroll
        for (int r = 0; r < 64 / RPI; ++r) {
            const int krow = r * RPI + lr;
            if (krow < n) {
                double *p = out + base + lf + (long long)krow * stride;
                if (W == 2) *reinterpret_cast<double2 *>(p) = make_double2(v[r][0] + 1.0, v[r][W - 1] + 1.0);
                else *p = v[r][0] + 1.0;
            }
        }
    }
}
__global__ void kcopy(const double2 *in, double2 *out, long long n2) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        double2 x = in[i]; x.x += 1.0; x.y += 1.0; out[i] = x;
    }
}
template <class Fn> static void timeit(const char *name, double bytes, Fn f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf(" \"%s\": {\"us\": %.1f, \"GBps\": %.0f},\n", name, best * 1e3, bytes / best * 1e-6);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int n = 56; const long long N = 9834496, stride = 3136, nfib = N / n, inner = 3136;
    double *x, *y; cudaMalloc(&x, N * 8); cudaMalloc(&y, N * 8); cudaMemset(x, 0, N * 8);
    const double bytes = 2.0 * N * 8;
    const int g = p.multiProcessorCount * 2;
    printf("{\n");
    timeit("F8_8B_lanes_64B_pieces(shipped pattern)", bytes, [&] { k<8, 1><<<g, 256>>>(x, y, nfib, n, stride, inner); });
    timeit("F16_8B_lanes_128B_pieces", bytes, [&] { k<16, 1><<<g, 256>>>(x, y, nfib, n, stride, inner); });
    timeit("F16_16B_lanes_128B_pieces", bytes, [&] { k<16, 2><<<g, 256>>>(x, y, nfib, n, stride, inner); });
    timeit("F32_8B_lanes_256B_pieces", bytes, [&] { k<32, 1><<<g, 256>>>(x, y, nfib, n, stride, inner); });
    timeit("F32_16B_lanes_256B_pieces", bytes, [&] { k<32, 2><<<g, 256>>>(x, y, nfib, n, stride, inner); });
    timeit("coalesced_copy_16B", bytes, [&] { kcopy<<<g * 4, 256>>>((const double2 *)x, (double2 *)y, N / 2); });
    printf(" \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
