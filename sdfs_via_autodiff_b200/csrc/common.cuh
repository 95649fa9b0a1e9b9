// Shared host/device definitions for libsdfs_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include "../../include/sdfs_b200.h"

namespace cg = cooperative_groups;

#define SDFS_MAX_RANKS 8
#define SDFS_THREADS 288               // 8 consumer warps + 1 TMA producer warp
#define SDFS_WARPS (SDFS_THREADS / 32)
#define SDFS_MAX_GRID 1024             // upper bound on cooperative grid size (slots)

struct sdfs_comm_state;                // comm.cu

struct sdfs_ctx {
    int device = 0;
    int sm_count = 0;
    int coop_supported = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    int64_t launches = 0;
    // optional per-launch CUDA-event timing of the dominant (dense pass) kernel
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;   // pairs (start, stop)
    size_t prof_used = 0;
    // multi-GPU
    int rank = 0, nranks = 1;
    sdfs_comm_state *comm = nullptr;
    // scratch for host-visible results of device loops
    void *d_status = nullptr;          // LoopStatus
    void *h_status = nullptr;          // pinned mirror
};

extern thread_local std::string g_last_error;

int sdfs_set_error(sdfs_ctx *ctx, int code, const char *fmt, ...);

#define CUDA_TRY(ctx, expr)                                                             \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            (void)cudaGetLastError(); /* clear the non-sticky error for later calls */  \
            return sdfs_set_error((ctx), _e == cudaErrorMemoryAllocation ? SDFS_ERR_NOMEM \
                                                                         : SDFS_ERR_CUDA, \
                                  "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,         \
                                  cudaGetErrorString(_e));                              \
        }                                                                               \
    } while (0)

#define ARG_CHECK(ctx, cond)                                                            \
    do {                                                                                \
        if (!(cond))                                                                    \
            return sdfs_set_error((ctx), SDFS_ERR_ARG, "%s:%d: argument check failed: %s", \
                                  __FILE__, __LINE__, #cond);                           \
    } while (0)

// pinned host word a kernel sets when a peer rank never reaches a fused exchange (checked at every sync)
static inline long long *ctx_h_abort(sdfs_ctx *ctx) { return (long long *)((char *)ctx->h_status + 4096 - 64); }
static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------
// Operator description passed by value to kernels.
// ---------------------------------------------------------------------------
struct DenseView {
    CUtensorMap tm;        // 2-D TMA descriptor of the local P slice (box 8 rows x 256 cols); valid iff vec2
    const double *P;       // rows [row_begin,row_end), leading dimension ld
    int64_t N, ld;
    int64_t row_begin, row_end;
    const double *a_row;   // N
    const double *a_col;   // N
    const double *e_sdf;   // N or null
    double beta, theta;
    int vec2;              // 1 when every row of P is 16-byte aligned (TMA path)
};

// Factor-structured operator: out = M_D ... M_1 applied mode by mode.
#define SDFS_MAX_DIMS 6
struct KronMode {
    const double *mat;     // [n_mats][n][n] row = current index, col = next index
    int dim;               // tensor axis contracted by this mode
    int mstride[SDFS_MAX_DIMS];  // matrix id = sum_d coord_d * mstride[d]
    // work decomposition of the fibre kernel (filled by kron_plan): the axes other than `dim`
    // split into "matrix" axes (mstride != 0; fixed per work item, so one factor matrix per item)
    // and "free" axes (enumerated by the threads, innermost fastest)
    int nF, nM;
    int Fshape[SDFS_MAX_DIMS], Mshape[SDFS_MAX_DIMS], Mmat[SDFS_MAX_DIMS];
    long long Fstride[SDFS_MAX_DIMS], Mstride[SDFS_MAX_DIMS];
    long long stride, Fcount, Mcount;
    // slab-sharded views (one process per GPU, leading axis split over the ranks; kron_restrict_leading):
    // the leading mode forms only the output rows [out0, out0 + nout) of its axis from the full input, the
    // other modes run on the local slab (axis 0 restricted in Fshape, element offset base_off)
    int out0, nout;
    long long base_off;
    // optional scaling of the matrix columns (next-period index), applied while the matrix is staged: the fused
    // apply folds a_col (a function of this axis alone) into the first contraction instead of reading an N-vector
    const double *colscale;
};
struct KronView {
    int D;
    int shape[SDFS_MAX_DIMS];
    int64_t N;
    int n_modes;
    KronMode modes[SDFS_MAX_DIMS];
    const double *a_row, *a_col, *e_sdf;
    double beta, theta;
    int lead0, leadn;                  // slab [lead0, lead0 + leadn) of axis 0 owned by this rank (whole axis when not sharded)
    int64_t row_begin, row_end;        // = lead0 * N / shape[0], (lead0 + leadn) * N / shape[0]
};

// Continuous-state operator (ssy/continuous_junnan/ssy_wc_ratio_continuous.py,
// gcy/continuous/gcy_wc_ratio_continuous.py): uniform interpolation grids, shock nodes/weights.
#define SDFS_STORAGE_CONT 3
struct ContView {
    int model, D, Q;
    int n[SDFS_MAX_DIMS];
    const double *grid[SDFS_MAX_DIMS];   // device: grid values per axis
    double g0[SDFS_MAX_DIMS], intv[SDFS_MAX_DIMS];   // grid[0] and grid[1]-grid[0] (utils.py:6-14)
    const double *nodes;                 // [D][Q] shocks (quadrature nodes or Monte-Carlo draws)
    const double *weights;               // [Q]
    double p[18];                        // model parameters, reference order
    double beta, gamma, theta, mu_c, phi_c;
    int64_t N;
    int64_t row_begin, row_end;          // states owned by this rank
};

struct sdfs_factors {
    sdfs_ctx *ctx = nullptr;
    int model = 0;
    int D = 0;
    int shapes[SDFS_MAX_DIMS] = {0};
    double params[18] = {0};
    int n_arrays = 0;
    double *d_arr[16] = {nullptr};
    int64_t n_elems[16] = {0};
};

struct sdfs_op {
    sdfs_ctx *ctx = nullptr;
    int storage = SDFS_STORAGE_DENSE;
    DenseView dv{};
    KronView kv{};                     // whole operator (builder kernels, dense expansion, sweeps)
    KronView kvs{};                    // what applications and solver loops of THIS rank contract: kv, or its slab-sharded restriction
    ContView cv{};
    double *cont_mem = nullptr;        // grids + nodes + weights of a continuous-state operator
    sdfs_factors *factors = nullptr;   // borrowed (kept alive by the host wrapper)
    // owned device memory
    double *own_P = nullptr, *own_a_row = nullptr, *own_a_col = nullptr, *own_e_sdf = nullptr;
    // work vectors (each ldv doubles, zero padded), allocated on first use
    int64_t ldv = 0;
    double *work = nullptr;            // NWORK * ldv
    int n_work = 0;
    double *slots = nullptr;           // reduction slots (single-GPU arena)
    double *kron_tmp[2] = {nullptr, nullptr};
    double gamma = 0, psi = 0, mu_c = 0;  // remembered for set_preferences
    int sweep_form = 0;                // SDFS_SWEEP_DENSE | SDFS_SWEEP_FACTOR
    bool kron_sharded = false;         // factor form with the leading axis split into per-rank slabs
    double *a_col_lead = nullptr;      // a_col along the axis of the first contraction (it is constant along the others)
    double *dense_tail = nullptr;      // DenseTail scratch of the dense application (rowdot.cuh): segment partials, then counters
};

int op_ensure_work(sdfs_op *op, int n_vectors);
void op_sync_kvs(sdfs_op *op);                                           // ops.cu
void op_rank_rows(const sdfs_op *op, int r, int64_t *rb, int64_t *re);   // ops.cu
bool op_is_sharded(const sdfs_op *op);                                   // ops.cu
int op_allgather(sdfs_op *op, double *d_vec);                            // ops.cu
static inline int64_t op_N(const sdfs_op *op) {
    return op->storage == SDFS_STORAGE_DENSE ? op->dv.N : (op->storage == SDFS_STORAGE_KRON ? op->kv.N : op->cv.N);
}

// ---------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// NaN-propagating max (jnp.max semantics, solvers.py:36)
__device__ __forceinline__ double nanmax(double a, double b) {
    return (a != a) ? a : ((b != b) ? b : fmax(a, b));
}
__device__ __forceinline__ double warp_nanmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Streaming 16-byte load of two P entries: read once, evict first, keep L1 for x.
__device__ __forceinline__ double2 ld_stream2(const double *p) {
    double2 r;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream1(const double *p) {
    double r;
    asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// x^e for x > 0 through exp(e log x): about half the latency of pow() (tools/lat_probe.cu: 443 vs
// 844 cycles).  Relative error <= ~|e log x| ulp (<= 1e-13 for the calibrations used), which the
// 1/theta power of T shrinks again by |theta|; NaN for x < 0 and inf for x = 0, e < 0 like pow.
__device__ __forceinline__ double pow_pos_lib(double x, double e) { return exp(e * log(x)); }

// Table-driven form of the same x^e = exp(e log x) (SDFS_FAST_POW, default on): ~26 fp64 instructions and three
// L1-resident table gathers instead of ~50 fp64 instructions.  The operator's prologue and epilogue are
// fp64-PIPE bound (DMMA shares that pipe, profiles/r02_pipe_probe.md), so this is where a factor-form
// application can still get cheaper.
//   log x:  x = 2^k m, m in [1,2); i = top 7 mantissa bits; r = m rc[i] - 1 (one FMA, |r| <= 2^-8);
//           log x = k ln2 + lc[i] + log1p(r), degree-7 polynomial        (rc[i] ~ 1/c_i, lc[i] = -log rc[i] exactly paired)
//   exp y:  n = rint(64 y / ln2), r = y - n ln2/64 (two FMAs), exp y = 2^(n>>6) e2[n & 63] (1 + r + ... + r^6/720)
// Validated on the host against long double (tools/gen_fastpow_tables.c + profiles/r02_fastpow.md): exp max relative
// error 1.9e-16, log max absolute error ~1 ulp of the result, x^e over w in [1, 2000], e in [-50, -5]: 7.3e-14
// against 5.0e-14 for libm's exp(e log x) - the |e| ulp(log x) term dominates both.  Zero, negative, subnormal,
// infinite and NaN arguments and results outside the normal range take the library path (NaN for x < 0, inf for
// x = 0 and e < 0, as the solver loops' NaN semantics require).
#ifndef SDFS_FAST_POW
#define SDFS_FAST_POW 1
#endif
#if SDFS_FAST_POW
#include "fastpow_tables.cuh"
__device__ __forceinline__ double pow_pos(double x, double e) {
    const unsigned hx = (unsigned)__double2hiint(x);
    if (hx - 0x00100000u >= 0x7fe00000u) return pow_pos_lib(x, e);          // not a positive normal number
    const int i = (hx >> 13) & 127;
    const double m = __hiloint2double((int)((hx & 0x000fffffu) | 0x3ff00000u), __double2loint(x));
    const double r = fma(m, __ldg(g_fp_rc + i), -1.0);
    double p = fma(r, 1.0 / 7, -1.0 / 6);
    p = fma(r, p, 1.0 / 5);
    p = fma(r, p, -1.0 / 4);
    p = fma(r, p, 1.0 / 3);
    p = fma(r, p, -0.5);
    const double kf = (double)((int)(hx >> 20) - 1023);
    const double hi = fma(kf, 0x1.62e42fefa38p-1, __ldg(g_fp_lc + i));       // ln2 high part: 32 trailing zero bits, the product is exact
    const double lg = hi + (r + fma(r * r, p, kf * 0x1.ef35793c7673p-45));
    const double y = e * lg;
    if (!(fabs(y) < 700.0)) return exp(y);                                    // over/underflow range, NaN
    const double t = fma(y, 0x1.71547652b82fep+6, 6755399441055744.0);        // 64 / ln2; magic rounding constant 1.5 * 2^52
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double q = fma(nf, -0x1.62e42fefa38p-7, y);
    q = fma(nf, -0x1.ef35793c7673p-51, q);
    double s = fma(q, 1.0 / 720, 1.0 / 120);
    s = fma(q, s, 1.0 / 24);
    s = fma(q, s, 1.0 / 6);
    s = fma(q, s, 0.5);
    const double em1 = fma(q * q, s, q);
    const double tj = __ldg(g_fp_e2 + (n & 63));
    const double v = fma(tj, em1, tj);
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}
// exp(e (log x + off)): the sweep's log-domain epilogue / prologue (log a_row resp. h_lambda is the offset) on the same tables
__device__ __forceinline__ double pow_pos_off(double x, double e, double off) {
    const unsigned hx = (unsigned)__double2hiint(x);
    if (hx - 0x00100000u >= 0x7fe00000u) return exp(e * (log(x) + off));   // not a positive normal number
    const int i = (hx >> 13) & 127;
    const double m = __hiloint2double((int)((hx & 0x000fffffu) | 0x3ff00000u), __double2loint(x));
    const double r = fma(m, __ldg(g_fp_rc + i), -1.0);
    double p = fma(r, 1.0 / 7, -1.0 / 6);
    p = fma(r, p, 1.0 / 5);
    p = fma(r, p, -1.0 / 4);
    p = fma(r, p, 1.0 / 3);
    p = fma(r, p, -0.5);
    const double kf = (double)((int)(hx >> 20) - 1023);
    const double hi = fma(kf, 0x1.62e42fefa38p-1, __ldg(g_fp_lc + i));       // ln2 high part: 32 trailing zero bits, the product is exact
    const double lg = hi + (r + fma(r * r, p, kf * 0x1.ef35793c7673p-45));
    const double y = e * (lg + off);
    if (!(fabs(y) < 700.0)) return exp(y);                                    // over/underflow range, NaN
    const double t = fma(y, 0x1.71547652b82fep+6, 6755399441055744.0);        // 64 / ln2; magic rounding constant 1.5 * 2^52
    const int n = __double2loint(t);
    const double nf = t - 6755399441055744.0;
    double q = fma(nf, -0x1.62e42fefa38p-7, y);
    q = fma(nf, -0x1.ef35793c7673p-51, q);
    double s = fma(q, 1.0 / 720, 1.0 / 120);
    s = fma(q, s, 1.0 / 24);
    s = fma(q, s, 1.0 / 6);
    s = fma(q, s, 0.5);
    const double em1 = fma(q * q, s, q);
    const double tj = __ldg(g_fp_e2 + (n & 63));
    const double v = fma(tj, em1, tj);
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}
#else
__device__ __forceinline__ double pow_pos(double x, double e) { return pow_pos_lib(x, e); }
__device__ __forceinline__ double pow_pos_off(double x, double e, double off) { return exp(e * (log(x) + off)); }
#endif

// fp64 tensor-core tile: D(8x8) += A(8x4, row) B(4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4]
// and D[l/4][2(l%4) + {0,1}].  (tcgen05 has no f64 kind; DMMA is the fp64 tensor path on sm_100a.)
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Coherent (weak, L1-cacheable) loads of the matvec input x.  x is rewritten between
// grid barriers inside the persistent loop kernels, so it must never be turned
// into a non-coherent ld.global.nc by the compiler.
__device__ __forceinline__ double2 ld_x2(const double *p) {
    double2 r;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ double ld_x1(const double *p) {
    double r;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}

#endif  // __CUDACC__
