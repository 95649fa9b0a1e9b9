// Access-pattern probe for the dense fp64 row-stream (y = P x) at large footprints.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/bw_probe tools/bw_probe.cu
// Run  :  tools/bw_probe <N> [reps]
// Variants: V0 blocked per-warp rows | V1 interleaved per-warp rows | V2 CTA-cooperative
// rows (LDG) | V3 pure read | V4 CTA-cooperative rows fed by cp.async.bulk (TMA) + mbarrier ring
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <cuda.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double2 ld_cs2(const double *p) {
    double2 r;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_nc2(const double *p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_x2(const double *p) {
    double2 r;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void k_fill(double *P, int64_t N, int64_t ld) {
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

// ---- per-warp rows (R rows, U chunks), ld % 64 == 0, columns padded with zeros ----
template <int R, int U, int LDK>
__device__ __forceinline__ void warp_rows(const double *p0, int64_t ld, const double *x, double (&out)[R]) {
    const int lane = threadIdx.x & 31;
    double a[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) a[r][0] = a[r][1] = 0.0;
    const int64_t nch = ld >> 6;
    int64_t c = 0;
    for (; c + U <= nch; c += U) {
        double2 pv[U][R], xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double *p = p0 + (int64_t)r * ld + ((c + u) << 6) + 2 * lane;
                pv[u][r] = LDK ? ld_nc2(p) : ld_cs2(p);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] = ld_x2(x + ((c + u) << 6) + 2 * lane);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                a[r][0] = fma(pv[u][r].x, xv[u].x, a[r][0]);
                a[r][1] = fma(pv[u][r].y, xv[u].y, a[r][1]);
            }
    }
    for (; c < nch; ++c) {
        const double2 xv = ld_x2(x + (c << 6) + 2 * lane);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double2 pv = ld_cs2(p0 + (int64_t)r * ld + (c << 6) + 2 * lane);
            a[r][0] = fma(pv.x, xv.x, a[r][0]);
            a[r][1] = fma(pv.y, xv.y, a[r][1]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out[r] = warp_sum(a[r][0] + a[r][1]);
}

template <int R, int U, int LDK, bool INTERLEAVED>
__global__ void __launch_bounds__(256, 2) k_warp_rows(const double *P, int64_t N, int64_t ld, const double *x, double *y) {
    const int lane = threadIdx.x & 31;
    const int wg = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
    const int64_t ngroups = N / R;   // probe: N % R == 0 assumed
    if (INTERLEAVED) {
        for (int64_t g = wg; g < ngroups; g += nw) {
            double s[R];
            warp_rows<R, U, LDK>(P + g * R * ld, ld, x, s);
            if (lane == 0)
#pragma unroll
                for (int r = 0; r < R; ++r) y[g * R + r] = s[r];
        }
    } else {
        const int64_t g0 = ngroups * wg / nw, g1 = ngroups * (wg + 1) / nw;
        for (int64_t g = g0; g < g1; ++g) {
            double s[R];
            warp_rows<R, U, LDK>(P + g * R * ld, ld, x, s);
            if (lane == 0)
#pragma unroll
                for (int r = 0; r < R; ++r) y[g * R + r] = s[r];
        }
    }
}

// ---- CTA-cooperative rows: the CTA's 8 warps split the columns of R rows ----------
// warp w handles chunks c = w, w+8, ... ; one step of the CTA covers 8 consecutive chunks = 4 KB per row
template <int R, int U, int LDK, int MINB>
__global__ void __launch_bounds__(256, MINB) k_cta_rows(const double *P, int64_t N, int64_t ld, const double *x, double *y) {
    __shared__ double part[8][R];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ngroups = N / R;
    const int64_t nch = ld >> 6;
    for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const double *p0 = P + g * R * ld;
        double a[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r][0] = a[r][1] = 0.0;
        int64_t c = warp;
        for (; c + 8 * (U - 1) < nch; c += 8 * U) {
            double2 pv[U][R], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double *p = p0 + (int64_t)r * ld + ((c + 8 * u) << 6) + 2 * lane;
                    pv[u][r] = LDK ? ld_nc2(p) : ld_cs2(p);
                }
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = ld_x2(x + ((c + 8 * u) << 6) + 2 * lane);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    a[r][0] = fma(pv[u][r].x, xv[u].x, a[r][0]);
                    a[r][1] = fma(pv[u][r].y, xv[u].y, a[r][1]);
                }
        }
        for (; c < nch; c += 8) {
            const double2 xv = ld_x2(x + (c << 6) + 2 * lane);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double2 pv = ld_cs2(p0 + (int64_t)r * ld + (c << 6) + 2 * lane);
                a[r][0] = fma(pv.x, xv.x, a[r][0]);
                a[r][1] = fma(pv.y, xv.y, a[r][1]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double s = warp_sum(a[r][0] + a[r][1]);
            if (lane == 0) part[warp][r] = s;
        }
        __syncthreads();
        if (threadIdx.x < R) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
            y[g * R + threadIdx.x] = s;
        }
        __syncthreads();
    }
}

// ---- pure read: upper bound for a streaming read of this footprint ------------------
__global__ void __launch_bounds__(256, 2) k_read(const double *P, int64_t total, double *sink) {
    double a0 = 0, a1 = 0;
    const int64_t n2 = total / 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t st = (int64_t)gridDim.x * blockDim.x;
    for (; i + 3 * st < n2; i += 4 * st) {
        double2 v0 = ld_cs2(P + 2 * i), v1 = ld_cs2(P + 2 * (i + st)), v2 = ld_cs2(P + 2 * (i + 2 * st)), v3 = ld_cs2(P + 2 * (i + 3 * st));
        a0 += v0.x + v1.x + v2.x + v3.x;
        a1 += v0.y + v1.y + v2.y + v3.y;
    }
    for (; i < n2; i += st) { double2 v = ld_cs2(P + 2 * i); a0 += v.x; a1 += v.y; }
    if (a0 + a1 == 123.456) sink[0] = a0;
}

// ---- TMA (cp.async.bulk) fed CTA-cooperative rows ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// stage = R rows x CW columns of P (+ the x chunk); 8 consumer warps + 1 producer warp
template <int R, int CW, int STAGES>
__global__ void __launch_bounds__(288, 1) k_tma_rows(const double *P, int64_t N, int64_t ld, const double *x, double *y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sP = reinterpret_cast<double *>(smem_raw);                 // [STAGES][R][CW]
    double *sX = sP + (size_t)STAGES * R * CW;                         // [STAGES][CW]
    uint64_t *full = reinterpret_cast<uint64_t *>(sX + (size_t)STAGES * CW);
    uint64_t *empty = full + STAGES;
    __shared__ double part[8][R];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ngroups = N / R;
    const int64_t nck = (ld + CW - 1) / CW;       // column blocks per row
    if (warp == 8) {
        // producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
                const double *p0 = P + g * R * ld;
                for (int64_t cb = 0; cb < nck; ++cb) {
                    const int64_t col = cb * CW;
                    const uint32_t w = (uint32_t)((ld - col < CW ? ld - col : CW) * 8);
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], w * (R + 1));
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        bulk_g2s(sP + ((size_t)stage * R + r) * CW, p0 + (int64_t)r * ld + col, w, &full[stage]);
                    bulk_g2s(sX + (size_t)stage * CW, x + col, w, &full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // consumers: warp w takes the 64-column slices w, w+8, ... of each stage for all R rows
        int stage = 0; uint32_t phase = 0;
        for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
            double a[R][2];
#pragma unroll
            for (int r = 0; r < R; ++r) a[r][0] = a[r][1] = 0.0;
            for (int64_t cb = 0; cb < nck; ++cb) {
                const int64_t col = cb * CW;
                const int wcols = (int)(ld - col < CW ? ld - col : CW);
                mbar_wait(&full[stage], phase);
                for (int cc = warp * 64; cc < wcols; cc += 8 * 64) {
                    const double2 xv = *reinterpret_cast<const double2 *>(sX + (size_t)stage * CW + cc + 2 * lane);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const double2 pv = *reinterpret_cast<const double2 *>(sP + ((size_t)stage * R + r) * CW + cc + 2 * lane);
                        a[r][0] = fma(pv.x, xv.x, a[r][0]);
                        a[r][1] = fma(pv.y, xv.y, a[r][1]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double s = warp_sum(a[r][0] + a[r][1]);
                if (lane == 0) part[warp][r] = s;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x < R) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
                y[g * R + threadIdx.x] = s;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
}

// ---- V5: 2-D tensor-map TMA, one box per stage, ring slot w owned by consumer warp w --------
__device__ __forceinline__ void tma_2d_g2s(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// stage = BR rows x BC cols (BC <= 256), NBOX boxes side by side along the columns
template <int BR, int BC, int NBOX>
__global__ void __launch_bounds__(288, 1) k_tma2d_rows(const __grid_constant__ CUtensorMap tm, int64_t N, int64_t ld,
                                                       const double *x, double *y) {
    constexpr int CW = BC * NBOX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sP = reinterpret_cast<double *>(smem_raw);                 // [8][NBOX][BR][BC]
    double *sX = sP + (size_t)8 * BR * CW;                             // [8][CW]
    uint64_t *full = reinterpret_cast<uint64_t *>(sX + (size_t)8 * CW);
    uint64_t *empty = full + 8;
    __shared__ double part[2][8][BR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ngroups = (N + BR - 1) / BR;
    const uint32_t nck = (uint32_t)((N + CW - 1) / CW);
    if (warp == 8) {
        if (lane == 0) {
            uint32_t t = 0;
            for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
                const int row0 = (int)(g * BR);
                for (uint32_t cb = 0; cb < nck; ++cb, ++t) {
                    const int slot = t & 7;
                    const int col = (int)(cb * CW);
                    const int64_t left = ld - col;
                    const uint32_t xbytes = (uint32_t)((left < CW ? left : CW) * 8);
                    mbar_wait(&empty[slot], ((t >> 3) & 1) ^ 1);
                    mbar_expect_tx(&full[slot], (uint32_t)(BR * CW * 8) + xbytes);
#pragma unroll
                    for (int b = 0; b < NBOX; ++b)
                        tma_2d_g2s(sP + ((size_t)slot * NBOX + b) * BR * BC, &tm, col + b * BC, row0, &full[slot]);
                    bulk_g2s(sX + (size_t)slot * CW, x + col, xbytes, &full[slot]);
                }
            }
        }
    } else {
        uint32_t t0 = 0;
        int buf = 0;
        const double *sp = sP + (size_t)warp * BR * CW;
        const double *sx = sX + (size_t)warp * CW;
        for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x, t0 += nck, buf ^= 1) {
            double a[BR][2];
#pragma unroll
            for (int r = 0; r < BR; ++r) a[r][0] = a[r][1] = 0.0;
            for (uint32_t cb = (uint32_t)(warp - (int)t0) & 7; cb < nck; cb += 8) {
                const uint32_t t = t0 + cb;
                mbar_wait(&full[warp], (t >> 3) & 1);
#pragma unroll
                for (int b = 0; b < NBOX; ++b)
#pragma unroll
                    for (int k = 0; k < BC / 64; ++k) {
                        const int cc = 64 * k + 2 * lane;
                        double2 xv = *reinterpret_cast<const double2 *>(sx + b * BC + cc);
                        if (cb + 1 == nck) {   // x beyond N is not zero-filled by the bulk copy: P is (tensor OOB)
                            const int64_t gc = (int64_t)cb * CW + b * BC + cc;
                            if (gc >= N) xv.x = 0.0;
                            if (gc + 1 >= N) xv.y = 0.0;
                        }
#pragma unroll
                        for (int r = 0; r < BR; ++r) {
                            const double2 pv = *reinterpret_cast<const double2 *>(sp + ((size_t)b * BR + r) * BC + cc);
                            a[r][0] = fma(pv.x, xv.x, a[r][0]);
                            a[r][1] = fma(pv.y, xv.y, a[r][1]);
                        }
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[warp]);
            }
#pragma unroll
            for (int r = 0; r < BR; ++r) {
                const double s = warp_sum(a[r][0] + a[r][1]);
                if (lane == 0) part[buf][warp][r] = s;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x < BR && g * BR + threadIdx.x < N) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += part[buf][w][threadIdx.x];
                y[g * BR + threadIdx.x] = s;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(const double *P, int64_t N, int64_t nrows, int64_t ld, int bc, int br, CUtensorMapL2promotion prom) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)nrows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)bc, (cuuint32_t)br};
    cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)P, dims, strides, box, es,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, prom,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(4); }
    return tm;
}

static double checksum(const std::vector<double> &v) { double s = 0; for (double x : v) s += x * 1e-6; return s; }

int main(int argc, char **argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 38416;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const int64_t ld = (N + 63) / 64 * 64;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    double *P, *x, *y, *yref;
    const size_t pbytes = (size_t)N * ld * 8;
    if (cudaMalloc(&P, pbytes) != cudaSuccess) { printf("cannot allocate %.1f GB\n", pbytes / 1e9); return 2; }
    CK(cudaMalloc(&x, (ld + 512) * 8)); CK(cudaMalloc(&y, N * 8)); CK(cudaMalloc(&yref, N * 8));
    k_fill<<<sms * 16, 256>>>(P, N, ld);
    k_fillx<<<64, 256>>>(x, N, ld + 512);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double gb = ((double)N * N * 8 + 2.0 * N * 8) / 1e9;
    std::vector<double> href(N), h(N);
    printf("N=%lld ld=%lld P=%.2f GB sms=%d reps=%d\n", (long long)N, (long long)ld, pbytes / 1e9, sms, reps);
    auto run = [&](const char *name, auto launch, bool is_ref, bool check) {
        CK(cudaMemset(y, 0, N * 8));
        launch(); launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) launch();
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-44s FAILED: %s\n", name, cudaGetErrorString(e)); exit(3); }
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        double maxrel = 0;
        if (check) {
            CK(cudaMemcpy(h.data(), y, N * 8, cudaMemcpyDeviceToHost));
            if (is_ref) href = h;
            for (int64_t i = 0; i < N; ++i) { double d = fabs(h[i] - href[i]) / (fabs(href[i]) + 1e-300); if (d > maxrel) maxrel = d; }
        }
        printf("%-44s %9.3f ms  %8.1f GB/s  maxrel %.2e\n", name, ms, gb / ms * 1e3, maxrel);
        fflush(stdout);
    };
    const int g2 = sms * 2;
    run("V3 pure read (cs loads)", [&] { k_read<<<g2 * 4, 256>>>(P, (int64_t)N * ld, yref); }, false, false);
    run("V0 warp rows blocked R4U4 cs", [&] { k_warp_rows<4, 4, 0, false><<<g2, 256>>>(P, N, ld, x, y); }, true, true);
    run("V1 warp rows interleaved R4U4 cs", [&] { k_warp_rows<4, 4, 0, true><<<g2, 256>>>(P, N, ld, x, y); }, false, true);
    run("V1 warp rows interleaved R2U8 cs", [&] { k_warp_rows<2, 8, 0, true><<<g2, 256>>>(P, N, ld, x, y); }, false, true);
    run("V1 warp rows interleaved R1U16 cs", [&] { k_warp_rows<1, 16, 0, true><<<g2, 256>>>(P, N, ld, x, y); }, false, true);
    run("V2 cta rows R4U4 cs 2cta/sm", [&] { k_cta_rows<4, 4, 0, 2><<<g2, 256>>>(P, N, ld, x, y); }, false, true);
    run("V2 cta rows R8U2 cs 2cta/sm", [&] { k_cta_rows<8, 2, 0, 2><<<g2, 256>>>(P, N, ld, x, y); }, false, true);
    run("V2 cta rows R8U2 nc 2cta/sm", [&] { k_cta_rows<8, 2, 1, 2><<<g2, 256>>>(P, N, ld, x, y); }, false, true);
    run("V2 cta rows R2U8 cs 2cta/sm", [&] { k_cta_rows<2, 8, 0, 2><<<g2, 256>>>(P, N, ld, x, y); }, false, true);
    run("V2 cta rows R8U4 cs 1cta/sm", [&] { k_cta_rows<8, 4, 0, 1><<<sms, 256>>>(P, N, ld, x, y); }, false, true);
    run("V2 cta rows R4U4 cs 4cta/sm(grid)", [&] { k_cta_rows<4, 4, 0, 2><<<sms * 4, 256>>>(P, N, ld, x, y); }, false, true);
    {
        constexpr int R = 8, CW = 256, ST = 8;
        const size_t sm = (size_t)ST * R * CW * 8 + (size_t)ST * CW * 8 + 2 * ST * 8;
        CK(cudaFuncSetAttribute(k_tma_rows<R, CW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        run("V4 tma rows R8 CW256 x8 stages", [&] { k_tma_rows<R, CW, ST><<<sms, 288, sm>>>(P, N, ld, x, y); }, false, true);
    }
    {
        constexpr int R = 4, CW = 512, ST = 8;
        const size_t sm = (size_t)ST * R * CW * 8 + (size_t)ST * CW * 8 + 2 * ST * 8;
        CK(cudaFuncSetAttribute(k_tma_rows<R, CW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        run("V4 tma rows R4 CW512 x8 stages", [&] { k_tma_rows<R, CW, ST><<<sms, 288, sm>>>(P, N, ld, x, y); }, false, true);
    }
    {
        constexpr int R = 8, CW = 512, ST = 5;
        const size_t sm = (size_t)ST * R * CW * 8 + (size_t)ST * CW * 8 + 2 * ST * 8;
        CK(cudaFuncSetAttribute(k_tma_rows<R, CW, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        run("V4 tma rows R8 CW512 x5 stages", [&] { k_tma_rows<R, CW, ST><<<sms, 288, sm>>>(P, N, ld, x, y); }, false, true);
    }
    {
        CUtensorMap tm = make_map(P, N, N, ld, 256, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        const size_t sm = (size_t)8 * 8 * 256 * 8 + (size_t)8 * 256 * 8 + 128;
        CK(cudaFuncSetAttribute(k_tma2d_rows<8, 256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        run("V5 tma2d box 8x256, slot/warp", [&] { k_tma2d_rows<8, 256, 1><<<sms, 288, sm>>>(tm, N, ld, x, y); }, false, true);
        CUtensorMap tm2 = make_map(P, N, N, ld, 256, 4, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
        const size_t sm2 = (size_t)8 * 4 * 512 * 8 + (size_t)8 * 512 * 8 + 128;
        CK(cudaFuncSetAttribute(k_tma2d_rows<4, 256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
        run("V6 tma2d 2 boxes 4x256, slot/warp", [&] { k_tma2d_rows<4, 256, 2><<<sms, 288, sm2>>>(tm2, N, ld, x, y); }, false, true);
        CUtensorMap tm3 = make_map(P, N, N, ld, 256, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        run("V5 tma2d box 8x256, no L2 promotion", [&] { k_tma2d_rows<8, 256, 1><<<sms, 288, sm>>>(tm3, N, ld, x, y); }, false, true);
    }
    printf("checksum %.6f\n", checksum(href));
    if (argc > 3) {
        // sustained mode: queue ~argv[3] seconds of launches, sample clocks/power mid-run
        const double secs = atof(argv[3]);
        auto sustained = [&](const char *name, auto launch) {
            launch(); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms1; cudaEventElapsedTime(&ms1, e0, e1);
            const int n = (int)(secs * 1e3 / ms1) + 1;
            cudaEventRecord(e0);
            for (int i = 0; i < n; ++i) launch();
            cudaEventRecord(e1);
            char buf[256] = {0};
            FILE *f = popen("sleep 0.6; nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits | head -1", "r");
            if (f) { if (!fgets(buf, sizeof(buf), f)) buf[0] = 0; pclose(f); }
            for (char *c = buf; *c; ++c) if (*c == '\n') *c = 0;
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= n;
            printf("SUSTAINED %-40s first %.3f ms | avg %.3f ms over %d launches  %8.1f GB/s | mid-run sm_mhz,power,pwrcap: %s\n",
                   name, ms1, ms, n, gb / ms * 1e3, buf);
            fflush(stdout);
        };
        constexpr int R = 4, CW = 512, ST = 8;
        const size_t sm = (size_t)ST * R * CW * 8 + (size_t)ST * CW * 8 + 2 * ST * 8;
        sustained("V4 tma R4 CW512 x8 grid=148", [&] { k_tma_rows<R, CW, ST><<<sms, 288, sm>>>(P, N, ld, x, y); });
        {
            CUtensorMap tm = make_map(P, N, N, ld, 256, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
            const size_t sm5 = (size_t)8 * 8 * 256 * 8 + (size_t)8 * 256 * 8 + 128;
            sustained("V5 tma2d box 8x256, slot/warp", [&] { k_tma2d_rows<8, 256, 1><<<sms, 288, sm5>>>(tm, N, ld, x, y); });
            CUtensorMap tm2 = make_map(P, N, N, ld, 256, 4, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
            const size_t sm6 = (size_t)8 * 4 * 512 * 8 + (size_t)8 * 512 * 8 + 128;
            sustained("V6 tma2d 2 boxes 4x256, slot/warp", [&] { k_tma2d_rows<4, 256, 2><<<sms, 288, sm6>>>(tm2, N, ld, x, y); });
        }
        sustained("V4 tma R4 CW512 x8 grid=111", [&] { k_tma_rows<R, CW, ST><<<111, 288, sm>>>(P, N, ld, x, y); });
        sustained("V4 tma R4 CW512 x8 grid=74", [&] { k_tma_rows<R, CW, ST><<<74, 288, sm>>>(P, N, ld, x, y); });
        sustained("V2 cta rows R4U4 cs 2cta/sm", [&] { k_cta_rows<4, 4, 0, 2><<<g2, 256>>>(P, N, ld, x, y); });
        sustained("V2 cta rows R8U4 cs 1cta/sm", [&] { k_cta_rows<8, 4, 0, 1><<<sms, 256>>>(P, N, ld, x, y); });
        sustained("V3 pure read", [&] { k_read<<<g2 * 4, 256>>>(P, (int64_t)N * ld, yref); });
        sustained("V4 tma R4 CW512 x8 grid=148 (again)", [&] { k_tma_rows<R, CW, ST><<<sms, 288, sm>>>(P, N, ld, x, y); });
    }
    return 0;
}
