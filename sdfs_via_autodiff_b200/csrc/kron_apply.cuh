// Factor-form (Kronecker) operator application as ONE cooperative kernel.
//
//   T w = 1 + beta (a_row . (M_D ... M_1)(a_col . w^theta))^(1/theta)      (ssy_wc_ratio.py:116-148,
//                                                                           gcy_wc_ratio.py:178-236)
// The w^theta prologue is evaluated inside the fragment loader of the first mode contraction (every
// element of w is loaded exactly once there; a_col, a function of that mode's axis alone in both models,
// is folded into the staged factor matrix), the ^(1/theta) epilogue inside the sink of the last one
// (every output is produced exactly once there); the modes in between ping-pong through two
// N-vectors that stay L2-resident up to ~8 M states, separated by grid barriers (1.2 us each,
// tools/pipe_probe.cu) instead of kernel boundaries.  One launch per application instead of
// prologue + n_modes + epilogue launches, and two N-vector round trips less.
// The transcendental work runs in rolled loops over a per-warp shared-memory stage (rowdot.cuh,
// kron_load_traits / kron_sink_traits): DMMA and the fp64 pipe are ONE pipe on B200
// (profiles/r02_pipe_probe.md), so nothing is gained by spreading pow over more warps - what matters is
// that the code stays compact and the registers few.
//
// Multi-GPU (one process per GPU): the view is restricted to this rank's slab of the leading axis
// (builder.cu::kron_restrict_leading).  The first contraction -- the leading axis -- reads the full input
// and forms only the slab's rows, everything after it is local, and the sink of the last contraction
// stores the finished rows straight into every rank's result buffer (NVLink peer stores, same flag trade
// as the dense pass), so when the kernel ends the whole vector is on every rank.
#pragma once
#include "common.cuh"
#include "rowdot.cuh"
#include "arena.cuh"

#define KRON_APPLY_THREADS 256
#define KRON_APPLY_SMEM ((KRON_SMAT_DOUBLES + (KRON_APPLY_THREADS / 32) * KRON_STAGE_DOUBLES_PER_WARP) * sizeof(double))

struct KronApplyArgs {
    int pmode;                  // 0: T   1: T + JVP (two contractions)   2: SDF (two)   3: plain P x
    const double *w, *v;        // operator input(s); v = direction (pmode 1) or x (pmode 3)
    double *tmp0, *tmp1;        // mode ping-pong (global element indices; a sharded rank touches its slab only)
    double *s0;                 // finished first contraction when two are needed
    unsigned long long *trace;  // optional: globaltimer at every phase boundary (SDFS_KRON_TRACE=1), else null
};

__device__ __forceinline__ void kron_trace(const KronApplyArgs &a, int slot) {
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[slot] = t;
    }
}

// element loader of the first contraction: the operator's prologue per element (a_col rides in the matrix)
struct KronLoadFused {
    int kind;                   // 0: the stored value (plain P x)   1: w^theta   2: w^(theta-1) v   3: w^(theta-1)
    const double *w, *v;
    double theta;
    __device__ __forceinline__ double raw(long long idx) const { return w[idx]; }
    __device__ __forceinline__ const double *ptr(long long idx) const { return w + idx; }
    __device__ __forceinline__ bool active() const { return kind != 0; }
    __device__ __forceinline__ double xform(long long idx, double x) const {
        if (kind == 1) return pow_pos(x, theta);
        const double t = pow_pos(x, theta - 1.0);
        return kind == 2 ? t * v[idx] : t;
    }
    __device__ __forceinline__ double operator()(long long idx) const { return kind ? xform(idx, w[idx]) : w[idx]; }
};
template <> struct kron_load_traits<KronLoadFused> { static constexpr bool xform = true; };

struct KronSinkStore {
    double *out;
    __device__ __forceinline__ void operator()(long long idx, double s) const { out[idx] = s; }
};

// sink of the last contraction: the operator's epilogue, called from the rolled stage loop.
//   pre(idx)   the per-row factor multiplied into the contraction before the loop (a_row; all loads up front)
//   quad(...)  up to four outputs per trip: rows idx + {0, 1, 8, 9} * stride, `left` = rows that exist from idx on
template <class Epilogue>
struct KronSinkEpi {
    const Epilogue &epi;
    const double *s0;           // first contraction's result when the epilogue needs two (else null)
    __device__ __forceinline__ double pre(long long idx) const { return epi.pre(idx); }
    __device__ __forceinline__ void operator()(long long idx, double s) const {      // non-staged paths: s unscaled
        const double f = epi.pre(idx);
        epi.one(idx, s0 ? f * s0[idx] : f * s, f * s);
    }
    __device__ __forceinline__ bool single_output() const { return epi.single_output(); }
    __device__ __forceinline__ void store1(long long idx, double v) const { epi.put(idx, v); }
    __device__ __forceinline__ void store2(long long idx, double v0, double v1) const { epi.put2(idx, v0, v1); }
    // single-output epilogues as values: a..d (pre-scaled contractions of rows idx + {0, 1, 8, 9} stride) are
    // replaced by the results; rows that do not exist (left <= 0, 1, 8, 9) get harmless inputs
    __device__ __forceinline__ void values(long long idx, long long stride, double &a, double &b, double &c, double &d, int left) const {
        const long long i1 = idx + stride, i2 = idx + 8 * stride, i3 = i2 + stride;
        double a0 = a, b0 = b, c0 = c, d0 = d;                 // first contraction (scaled) when two feed the epilogue
        if (s0) {
            a0 = left > 0 ? epi.pre(idx) * s0[idx] : 1.0; b0 = left > 1 ? epi.pre(i1) * s0[i1] : 1.0;
            c0 = left > 8 ? epi.pre(i2) * s0[i2] : 1.0; d0 = left > 9 ? epi.pre(i3) * s0[i3] : 1.0;
        }
        if (left <= 1) { b = 1.0; b0 = 1.0; }
        if (left <= 8) { c = 1.0; c0 = 1.0; }
        if (left <= 9) { d = 1.0; d0 = 1.0; }
        epi.values4(a0, a, b0, b, c0, c, d0, d);
    }
    // multi-output epilogues (SDF): store from here
    __device__ __forceinline__ void quad(long long idx, long long stride, double a, double b, double c, double d, int left) const {
        const long long i1 = idx + stride, i2 = idx + 8 * stride, i3 = i2 + stride;
        if (left > 0) epi.one(idx, s0 ? epi.pre(idx) * s0[idx] : a, a);
        if (left > 1) epi.one(i1, s0 ? epi.pre(i1) * s0[i1] : b, b);
        if (left > 8) epi.one(i2, s0 ? epi.pre(i2) * s0[i2] : c, c);
        if (left > 9) epi.one(i3, s0 ? epi.pre(i3) * s0[i3] : d, d);
    }
};
template <class E> struct kron_sink_traits<KronSinkEpi<E>> { static constexpr bool staged = true; };

// epilogue.pre(n) = a_row[n]; epilogue.one(n, a_row s0, a_row s1) / epilogue.four(...) for every row n of this rank
template <class Epilogue>
__device__ __forceinline__ void kron_apply_device(cg::grid_group &grid, const KronView &kv, const KronApplyArgs &a,
                                                  double *smat, double *stage, Epilogue &&epilogue) {
    const int nx = (a.pmode == 1 || a.pmode == 2) ? 2 : 1;
    const int last = kv.n_modes - 1;          // >= 1 (checked on the host)
    const KronShare share(stage);
    int slot = 0;
    kron_trace(a, slot++);
    for (int pass = 0; pass < nx; ++pass) {
        for (int m = 0; m <= last; ++m) {
            const double *in = (m & 1) ? a.tmp0 : a.tmp1;          // mode m - 1 wrote tmp[(m - 1) & 1]
            double *out = (m & 1) ? a.tmp1 : a.tmp0;
            if (m == 0) {
                // plain P x: the loader degenerates to the stored value (kind 0, w := x)
                KronLoadFused ld{a.pmode == 3 ? 0 : ((pass == 0) ? 1 : (a.pmode == 1 ? 2 : 3)), a.pmode == 3 ? a.v : a.w, a.v, kv.theta};
                kron_mode_apply_ld<true, false, true>(kv, 0, ld, smat, KronSinkStore{out}, share);
                grid.sync();
                kron_trace(a, slot++);
            } else if (m < last || (nx == 2 && pass == 0)) {
                kron_mode_apply_ld<true, false, false>(kv, m, KronLoadPlain{in}, smat, KronSinkStore{m < last ? out : a.s0}, share);
                grid.sync();
                kron_trace(a, slot++);
            } else {
                typedef typename kron_bare<Epilogue>::type EpiT;
                KronSinkEpi<EpiT> sink{epilogue, nx == 2 ? a.s0 : nullptr};
                kron_mode_apply_ld<true, false, false>(kv, m, KronLoadPlain{in}, smat, sink, share);
                if (a.trace) { grid.sync(); kron_trace(a, slot++); }
            }
        }
    }
}
