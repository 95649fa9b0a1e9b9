// Streaming row-tile dot products: the HBM-bound core of every dense operator
// application  (s = P x with fused prologue/epilogue).
//
// Primary path (rows 16-byte aligned): TMA-fed CTA-cooperative rows.
//   A CTA owns groups of TR = 8 consecutive rows (group g -> CTA g mod grid, so the
//   whole grid sweeps one compact window of P).  One producer thread streams each
//   group as [8 rows x 256 columns] boxes with ONE cp.async.bulk.tensor.2d (TMA
//   tensor copy, 16 KB) plus one bulk copy of the matching 2 KB chunk of x per stage
//   into a TST = 8 deep shared-memory ring guarded by full/empty mbarriers (128 KB
//   of P in flight per SM).  Ring slot w belongs to consumer warp w, which contracts the whole stage
//   (conflict-free 16-byte LDS), accumulates per-lane partials in a fixed column
//   order, and the warps' row partials are combined in a fixed order per row group.
//   The TMA ring measured 7.41 TB/s on an 88 GB P (profiles/r01_bw_probe.md), ahead
//   of every LDG variant.
// Fallback path (odd leading dimension / unaligned user P): per-warp rows with
//   8-byte streaming loads.
// A row's result is a fixed-order sum for a given (N, path, position of its group in the CTA's schedule): runs repeat
// bit for bit; across grid sizes and row shardings the order of the eight per-warp partials (and, for the groups of the
// last partial wave, the column segments - see DenseTail) differs, i.e. rows agree to rounding (~1e-16 relative).
#pragma once
#include "common.cuh"

#define TR 8          // rows per group = rows of one TMA box
#define TCW 256       // columns per stage = columns of one TMA box (hardware limit 256)
#define TST 8         // stages in the ring
#define CONSUMER_WARPS 8
#define PRODUCER_WARP 8

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
// 1-D bulk copy (x chunk) and 2-D tensor copy (P box), both completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_2d_g2s(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// Shared-memory ring.  NX = number of x vectors contracted in the same P stream.
// Ring slot w belongs to consumer warp w: the producer fills slots round-robin, warp w
// consumes every 8th stage on its own, so a warp has 8 stage-times to turn one stage
// around and a stage is released by a single arrive.
template <int NX>
struct RowPipe {
    double P[TST][TR][TCW];          // 128 KB: one 8 x 256 TMA box per slot
    double X[TST][NX][TCW];          // 16 KB per x vector
    double part[2][CONSUMER_WARPS][NX][TR];   // per-warp row partials, double buffered by group
    uint64_t full[TST], empty[TST];
};
struct PipeState {                   // stages issued/consumed by this CTA so far (persists across passes)
    uint32_t t;
};

template <int NX>
__device__ __forceinline__ void pipe_init(RowPipe<NX> *rp, PipeState &st) {
    static_assert(TST == CONSUMER_WARPS, "one ring slot per consumer warp");
    if (threadIdx.x == 0) {
        for (int s = 0; s < TST; ++s) {
            mbar_init(&rp->full[s], 1);
            mbar_init(&rp->empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    st.t = 0;
    __syncthreads();
}

// One pass over this rank's rows with the TMA ring.  Block = 9 warps (8 consumers +
// 1 producer).  Per stage the producer issues ONE 2-D tensor copy (8 rows x 256 columns of
// P = 16 KB; rows/columns outside the matrix are zero-filled by the TMA unit, so ragged
// edges and odd N need no special cases) and ONE bulk copy of the matching 2 KB of x.
// epi(n_global, s0, s1) runs on thread r (< rows in group) of warp 0.
// Preconditions: dv.vec2 (dv.tm valid), x0/x1 16-byte aligned, readable and finite up to
// index round_up(N, 2).  `dv` must live in kernel parameter space (__grid_constant__).
// Tail split (optional, `tail.buf` non-null): the groups beyond the last FULL wave of the grid -- R = ngroups mod grid
// of them -- keep R CTAs streaming while the rest of the chip idles.  When R is small (S = grid / R >= 8 segments) each of
// those groups is cut into S column segments, one per CTA; a segment's row partials go to `tail.buf`, and the CTA that
// arrives last at the group's counter adds the S segments in segment order and runs the epilogue: same result whichever
// CTA finishes last.  Measured on one rank's slab of the 88 GB operator (tools/dense_tail.py, profiles/r02_dense_tail.md):
// 13 122 rows (R = 13, S = 11) 1521 -> 1509 us, against 1505 us for a perfectly divisible slab.  A partial wave with many
// CTAs (R = 25, 49: 4 and 2 ranks) already streams at the full HBM rate and the split only adds its combine (+16 / +24 us),
// hence the threshold.
#define DENSE_TAIL_MIN_SPLIT 8
struct DenseTail {
    double *buf;            // [grid][2][TR] segment partials (slot = CTA)
    unsigned int *cnt;      // [grid] arrivals per tail group, zero between passes (the last arriver resets its counter)
};

template <int NX, class Epi>
__device__ __forceinline__ void dense_pass_tma(const DenseView &dv, const double *x0, const double *x1,
                                               RowPipe<NX> *rp, PipeState &st, Epi &&epi, DenseTail tail = DenseTail{nullptr, nullptr}) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nloc = dv.row_end - dv.row_begin;
    const int64_t ngroups = (nloc + TR - 1) / TR;
    const int64_t ncols = (dv.N + 1) & ~(int64_t)1;          // x columns to move (even count)
    const uint32_t nck = (uint32_t)((dv.N + TCW - 1) / TCW);
    const uint32_t t_begin = st.t;
    // whole groups: g = blockIdx.x + k gridDim.x below g_main; then at most one column segment of a tail group
    int64_t g_main = ngroups;
    int tail_r = 0, tail_s = 0;                               // tail groups, segments per tail group
    if (tail.buf) {
        const int r = (int)(ngroups % gridDim.x);
        if (r > 0 && ngroups > (int64_t)gridDim.x && (int)gridDim.x / r >= DENSE_TAIL_MIN_SPLIT && nck >= 64) {
            tail_r = r; tail_s = (int)gridDim.x / r; g_main = ngroups - r;
        }
    }
    const int64_t my_groups = (g_main > (int64_t)blockIdx.x) ? (g_main - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const bool has_seg = tail_r > 0 && (int)blockIdx.x < tail_r * tail_s;
    const int seg_g = has_seg ? (int)blockIdx.x % tail_r : 0, seg_s = has_seg ? (int)blockIdx.x / tail_r : 0;
    const uint32_t seg_cb0 = has_seg ? (uint32_t)((uint64_t)nck * seg_s / tail_s) : 0;
    const uint32_t seg_cb1 = has_seg ? (uint32_t)((uint64_t)nck * (seg_s + 1) / tail_s) : 0;
    const int64_t n_items = my_groups + (has_seg ? 1 : 0);
    // every thread advances the shared stage counter identically
    st.t = t_begin + (uint32_t)my_groups * nck + (seg_cb1 - seg_cb0);
    if (warp == PRODUCER_WARP) {
        if (lane == 0) {
            // order the generic-proxy stores that produced x (before the last grid barrier)
            // ahead of the async-proxy reads below
            asm volatile("fence.proxy.async;" ::: "memory");
            uint32_t t = t_begin;
            for (int64_t i = 0; i < n_items; ++i) {
                const bool seg = i >= my_groups;
                const int64_t g = seg ? g_main + seg_g : (int64_t)blockIdx.x + i * gridDim.x;
                const uint32_t cb0 = seg ? seg_cb0 : 0, cb1 = seg ? seg_cb1 : nck;
                const int row0 = (int)(g * TR);
                for (uint32_t cb = cb0; cb < cb1; ++cb, ++t) {
                    const int col = (int)(cb * TCW);
                    const int64_t left = ncols - col;
                    const uint32_t xbytes = (uint32_t)((left < TCW ? left : TCW) * 8);
                    const int slot = t & (TST - 1);
                    mbar_wait(&rp->empty[slot], ((t >> 3) & 1) ^ 1);
                    mbar_expect_tx(&rp->full[slot], (uint32_t)(TR * TCW * 8) + (uint32_t)NX * xbytes);
                    tma_2d_g2s(&rp->P[slot][0][0], &dv.tm, col, row0, &rp->full[slot]);
                    bulk_g2s(&rp->X[slot][0][0], x0 + col, xbytes, &rp->full[slot]);
                    if (NX > 1) bulk_g2s(&rp->X[slot][NX - 1][0], x1 + col, xbytes, &rp->full[slot]);
                }
            }
        }
    } else {
        uint32_t t0 = t_begin;                   // stage number of the first column block of the current item
        int buf = 0;
        const double *sp = &rp->P[warp][0][0];
        const double *sx = &rp->X[warp][0][0];
        for (int64_t i = 0; i < n_items; ++i, buf ^= 1) {
            const bool seg = i >= my_groups;
            const int64_t g = seg ? g_main + seg_g : (int64_t)blockIdx.x + i * gridDim.x;
            const uint32_t cb0 = seg ? seg_cb0 : 0, cb1 = seg ? seg_cb1 : nck;
            const int64_t r0 = g * TR;
            const int nr = (int)(nloc - r0 < TR ? nloc - r0 : TR);
            double a[NX][TR][2];
#pragma unroll
            for (int q = 0; q < NX; ++q)
#pragma unroll
                for (int r = 0; r < TR; ++r) a[q][r][0] = a[q][r][1] = 0.0;
            // this warp's column blocks: those whose stage number is = warp (mod 8)
            for (uint32_t cb = cb0 + ((uint32_t)(warp - (int)t0) & (TST - 1)); cb < cb1; cb += TST) {
                const uint32_t t = t0 + (cb - cb0);
                const bool last = (cb + 1 == nck);
                mbar_wait(&rp->full[warp], (t >> 3) & 1);
                // lane takes columns 64k + 2 lane, k = 0..3, of all TR rows (conflict-free LDS.128)
#pragma unroll
                for (int k = 0; k < TCW / 64; ++k) {
                    const int cc = 64 * k + 2 * lane;
                    double2 xv[NX];
                    xv[0] = *reinterpret_cast<const double2 *>(sx + cc);
                    if (NX > 1) xv[NX - 1] = *reinterpret_cast<const double2 *>(sx + (NX - 1) * TCW + cc);
                    if (last) {
                        // beyond N the P box is zero-filled by the TMA unit, but the x slot keeps
                        // stale shared memory there: mask it (0 * garbage must stay 0)
                        const int64_t gc = (int64_t)cb * TCW + cc;
                        if (gc >= dv.N) { xv[0].x = 0.0; if (NX > 1) xv[NX - 1].x = 0.0; }
                        if (gc + 1 >= dv.N) { xv[0].y = 0.0; if (NX > 1) xv[NX - 1].y = 0.0; }
                    }
#pragma unroll
                    for (int r = 0; r < TR; ++r) {
                        const double2 pv = *reinterpret_cast<const double2 *>(sp + r * TCW + cc);
#pragma unroll
                        for (int q = 0; q < NX; ++q) {
                            a[q][r][0] = fma(pv.x, xv[q].x, a[q][r][0]);
                            a[q][r][1] = fma(pv.y, xv[q].y, a[q][r][1]);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&rp->empty[warp]);
            }
            t0 += cb1 - cb0;
#pragma unroll
            for (int q = 0; q < NX; ++q)
#pragma unroll
                for (int r = 0; r < TR; ++r) {
                    const double s = warp_sum(a[q][r][0] + a[q][r][1]);
                    if (lane == 0) rp->part[buf][warp][q][r] = s;
                }
            // one barrier per group: part[] is double buffered, so the other warps run on into
            // the next group while warp 0's first lanes evaluate the epilogue (pow etc.)
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (!seg) {
                if (threadIdx.x < nr) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int w = 0; w < CONSUMER_WARPS; ++w) {
                        s0 += rp->part[buf][w][0][threadIdx.x];
                        if (NX > 1) s1 += rp->part[buf][w][NX - 1][threadIdx.x];
                    }
                    epi(dv.row_begin + r0 + threadIdx.x, s0, NX > 1 ? s1 : s0);
                }
            } else if (warp == 0) {
                // this CTA's segment of tail group seg_g: publish, count, and let the last arriver finish the rows
                if (lane < TR) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int w = 0; w < CONSUMER_WARPS; ++w) {
                        s0 += rp->part[buf][w][0][lane];
                        if (NX > 1) s1 += rp->part[buf][w][NX - 1][lane];
                    }
                    double *slot = tail.buf + ((size_t)blockIdx.x * 2) * TR;
                    __stcg(slot + lane, s0);
                    if (NX > 1) __stcg(slot + TR + lane, s1);
                    __threadfence();
                }
                __syncwarp();
                unsigned int prev = 0;
                if (lane == 0) prev = atomicAdd(tail.cnt + seg_g, 1u);
                prev = __shfl_sync(0xffffffffu, prev, 0);
                if (prev == (unsigned int)tail_s - 1) {
                    __threadfence();
                    if (lane < nr) {
                        double s0 = 0.0, s1 = 0.0;
                        for (int sgm = 0; sgm < tail_s; ++sgm) {               // segment order, whoever arrives last
                            const double *slot = tail.buf + ((size_t)(sgm * tail_r + seg_g) * 2) * TR;
                            s0 += __ldcg(slot + lane);
                            if (NX > 1) s1 += __ldcg(slot + TR + lane);
                        }
                        epi(dv.row_begin + r0 + lane, s0, NX > 1 ? s1 : s0);
                    }
                    if (lane == 0) tail.cnt[seg_g] = 0;                        // every segment has arrived: ready for the next pass
                }
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // part[] quiescent before the next pass
    }
}

// ---------------------------------------------------------------------------
// Fallback: per-warp rows with direct streaming loads (any alignment).
// ---------------------------------------------------------------------------
template <int R, int U, int NX>
__device__ __forceinline__ void warp_rows_dot(const double *__restrict__ p0, int64_t ld,
                                              const double *x0, const double *x1, int64_t N,
                                              double (&out)[NX][R]) {
    const int lane = threadIdx.x & 31;
    double a[NX][R][2];
#pragma unroll
    for (int q = 0; q < NX; ++q)
#pragma unroll
        for (int r = 0; r < R; ++r) a[q][r][0] = a[q][r][1] = 0.0;
    const double *pr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) pr[r] = p0 + (int64_t)r * ld + 2 * lane;
    const double *xa = x0 + 2 * lane;
    const double *xb = (NX > 1) ? (x1 + 2 * lane) : xa;
    const int64_t nfull = N >> 6;
    int64_t c = 0;
    for (; c + U <= nfull; c += U) {
        double2 pv[U][R];
        double2 xv[NX][U];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const double *p = pr[r] + ((c + u) << 6);
                pv[u][r] = make_double2(ld_stream1(p), ld_stream1(p + 1));
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xv[0][u] = make_double2(ld_x1(xa + ((c + u) << 6)), ld_x1(xa + ((c + u) << 6) + 1));
            if (NX > 1) xv[NX - 1][u] = make_double2(ld_x1(xb + ((c + u) << 6)), ld_x1(xb + ((c + u) << 6) + 1));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int q = 0; q < NX; ++q) {
                    a[q][r][0] = fma(pv[u][r].x, xv[q][u].x, a[q][r][0]);
                    a[q][r][1] = fma(pv[u][r].y, xv[q][u].y, a[q][r][1]);
                }
    }
    for (; c < nfull; ++c) {
        double2 xv[NX];
        xv[0] = make_double2(ld_x1(xa + (c << 6)), ld_x1(xa + (c << 6) + 1));
        if (NX > 1) xv[NX - 1] = make_double2(ld_x1(xb + (c << 6)), ld_x1(xb + (c << 6) + 1));
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double *p = pr[r] + (c << 6);
            const double2 pv = make_double2(ld_stream1(p), ld_stream1(p + 1));
#pragma unroll
            for (int q = 0; q < NX; ++q) {
                a[q][r][0] = fma(pv.x, xv[q].x, a[q][r][0]);
                a[q][r][1] = fma(pv.y, xv[q].y, a[q][r][1]);
            }
        }
    }
    // ragged tail: columns [64*nfull, N)
    const int64_t col = (nfull << 6) + 2 * lane;
    if (col < N) {
        const bool two = (col + 1 < N);
        double xs[NX][2];
        xs[0][0] = ld_x1(x0 + col);
        xs[0][1] = two ? ld_x1(x0 + col + 1) : 0.0;
        if (NX > 1) {
            xs[NX - 1][0] = ld_x1(x1 + col);
            xs[NX - 1][1] = two ? ld_x1(x1 + col + 1) : 0.0;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double *p = p0 + (int64_t)r * ld + col;
            const double q0 = ld_stream1(p);
            const double q1 = two ? ld_stream1(p + 1) : 0.0;
#pragma unroll
            for (int q = 0; q < NX; ++q) {
                a[q][r][0] = fma(q0, xs[q][0], a[q][r][0]);
                a[q][r][1] = fma(q1, xs[q][1], a[q][r][1]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NX; ++q)
#pragma unroll
        for (int r = 0; r < R; ++r) out[q][r] = warp_sum(a[q][r][0] + a[q][r][1]);
}

template <int R>
__device__ __forceinline__ double pick(const double (&v)[R], int lane) {
    double s = v[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
        if (lane == r) s = v[r];
    return s;
}

// Warp `wg` of `nw` owns row groups wg, wg+nw, ... (groups of 4 rows; the ragged last
// group is finished with 2- and 1-row passes).  epi runs on one lane per row.
template <int NX, class Epi>
__device__ __forceinline__ void dense_pass_ldg(const DenseView &dv, const double *x0, const double *x1,
                                               int wg, int nw, Epi &&epi) {
    const int lane = threadIdx.x & 31;
    const int64_t nloc = dv.row_end - dv.row_begin;
    const int64_t ngroups = (nloc + 3) / 4;
    for (int64_t g = wg; g < ngroups; g += nw) {
        int64_t r = g * 4;
        const int64_t r1 = (r + 4 < nloc) ? r + 4 : nloc;
        if (r + 4 <= r1) {
            double s[NX][4];
            warp_rows_dot<4, 4, NX>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
            if (lane < 4) epi(dv.row_begin + r + lane, pick<4>(s[0], lane), pick<4>(s[NX - 1], lane));
            continue;
        }
        if (r + 2 <= r1) {
            double s[NX][2];
            warp_rows_dot<2, 4, NX>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
            if (lane < 2) epi(dv.row_begin + r + lane, pick<2>(s[0], lane), pick<2>(s[NX - 1], lane));
            r += 2;
        }
        if (r < r1) {
            double s[NX][1];
            warp_rows_dot<1, 8, NX>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
            if (lane < 1) epi(dv.row_begin + r, s[0][0], s[NX - 1][0]);
        }
    }
}

// Dispatch on the operator's alignment class.
template <int NX, class Epi>
__device__ __forceinline__ void dense_pass(const DenseView &dv, const double *x0, const double *x1,
                                           RowPipe<NX> *rp, PipeState &st, Epi &&epi, DenseTail tail = DenseTail{nullptr, nullptr}) {
    if (dv.vec2) {
        dense_pass_tma<NX>(dv, x0, x1, rp, st, epi, tail);
    } else {
        const int wg = blockIdx.x * SDFS_WARPS + (threadIdx.x >> 5);
        dense_pass_ldg<NX>(dv, x0, x1, wg, gridDim.x * SDFS_WARPS, epi);
    }
}

// Factor-structured apply: one mode contraction
//   out[idx] = sum_j M[mat_id(idx)][i_m][j] * in[idx with i_m := j]
// over a grid-stride range of output elements.  Called once per mode with a grid
// barrier between calls; the last mode feeds the epilogue instead of storing.
// Default element loader of the contractions: the stored value.  The fused apply passes a loader that
// evaluates the operator's prologue (a_col w^theta ...) on the way in, so the first contraction reads w itself.
struct KronLoadPlain {
    const double *in;
    __device__ __forceinline__ double operator()(long long idx) const { return in[idx]; }
    __device__ __forceinline__ double raw(long long idx) const { return in[idx]; }
    __device__ __forceinline__ double xform(long long, double x) const { return x; }
    __device__ __forceinline__ const double *ptr(long long idx) const { return in + idx; }
};
// Hooks of the tensor-core contraction for loaders / sinks that carry transcendental work (the fused
// prologue and epilogue of the operator).  Evaluating a pow per fragment element inside the fully unrolled
// fragment code costs ~30 inlined pows per instantiation (20 KB of spills and minutes of ptxas time when it
// was tried).  Instead the warp parks the fragment values in a private shared-memory stage ([16][32]
// doubles per warp, an indexable register file: lane-private columns, no synchronisation) and a ROLLED loop
// transforms them two at a time (two independent log/exp chains) - compact code, few live registers.
//   loader with xform:  raw(idx) is the memory read (issued early: prefetch), xform(idx, x) the arithmetic
//   staged sink:        operator()(idx, s) is called from the rolled loop instead of the unrolled tail
template <class L> struct kron_load_traits { static constexpr bool xform = false; };
template <class S> struct kron_sink_traits { static constexpr bool staged = false; };
template <class T> struct kron_bare { typedef T type; };
template <class T> struct kron_bare<T &> { typedef T type; };
template <class T> struct kron_bare<const T &> { typedef T type; };
template <class T> struct kron_bare<T &&> { typedef T type; };
template <class T> struct kron_bare<const T> { typedef T type; };

template <class Load, class Sink>
__device__ __forceinline__ void kron_mode_pass(const KronView &kv, int m, Load &&load, int64_t tid,
                                               int64_t nthreads, Sink &&sink) {
    const KronMode &md = kv.modes[m];
    const int dim = md.dim;
    const int n = kv.shape[dim];
    const int64_t stride = md.stride;
    // output elements of this rank: every axis in full except the slab axis (axis 0), which a sharded view
    // restricts to [lead0, lead0 + leadn) (as output rows of the leading mode, as a free axis of the others)
    const int64_t inner = kv.N / kv.shape[0];
    const int64_t n_out = (int64_t)kv.leadn * inner, off = (int64_t)kv.lead0 * inner;
    for (int64_t e = tid; e < n_out; e += nthreads) {
        const int64_t idx = e + off;
        int64_t rem = idx;
        int mat = 0, im = 0;
        for (int d = kv.D - 1; d >= 0; --d) {
            const int cd = (int)(rem % kv.shape[d]);
            rem /= kv.shape[d];
            mat += cd * md.mstride[d];
            if (d == dim) im = cd;
        }
        const double *row = md.mat + ((int64_t)mat * n + im) * n;
        const int64_t src = idx - (int64_t)im * stride;
        double acc = 0.0;
        if (md.colscale) for (int j = 0; j < n; ++j) acc = fma(row[j] * md.colscale[j], load(src + (int64_t)j * stride), acc);
        else for (int j = 0; j < n; ++j) acc = fma(row[j], load(src + (int64_t)j * stride), acc);
        sink(idx, acc);
    }
}

// ---------------------------------------------------------------------------
// Register-tiled mode contraction ("fibre" kernel).
// A work item = (one combination of the matrix axes, a chunk of blockDim free-axis
// combinations): all its fibres use the same n x n factor matrix, which is staged in
// shared memory (rows zero-padded to NMAX) and read with broadcast LDS.128.  Each thread
// loads its fibre (n inputs, stride = stride of the contracted axis) into registers,
// forms the n outputs with n*NMAX FMAs and stores them (or feeds the epilogue).
// Per fibre: n loads + n stores + n^2 FMAs, against 2 n^2 cached loads in kron_mode_pass.
// Items are distributed round-robin over the CTAs (item -> CTA item mod grid), so the same
// function serves plain launches and the persistent loop kernels.
// ---------------------------------------------------------------------------
#define KRON_NMAX_LIMIT 64
// Which share of a mode's work items a CTA takes: the launch grid by default, {0, 1} when one CTA
// contracts a whole (shared-memory resident) vector by itself (fused sweep kernel).
struct KronShare {
    unsigned cta, nctas;
    double *stage;       // per-CTA staging area (KRON_STAGE_DOUBLES_PER_WARP per warp) when the loader / sink uses it
    __device__ __forceinline__ KronShare() : cta(blockIdx.x), nctas(gridDim.x), stage(nullptr) {}
    __device__ __forceinline__ KronShare(unsigned c, unsigned n) : cta(c), nctas(n), stage(nullptr) {}
    __device__ __forceinline__ explicit KronShare(double *st) : cta(blockIdx.x), nctas(gridDim.x), stage(st) {}
};
#define KRON_TC_MIN 9        // shortest axis contracted on the tensor cores (two 8-row output tiles);
                             // kron_mode_fibre<8> serves everything below, so this must stay <= 9
template <int NMAX, class Load, class Sink>
__device__ __forceinline__ void kron_mode_fibre(const KronView &kv, int m, Load &&load, double *smat /* n*NMAX */,
                                                Sink &&sink, KronShare share = KronShare()) {
    const KronMode &md = kv.modes[m];
    const int n = kv.shape[md.dim];
    const long long chunks = (md.Fcount + blockDim.x - 1) / blockDim.x;
    const long long items = md.Mcount * chunks;
    for (long long item = share.cta; item < items; item += share.nctas) {
        const long long mc = item / chunks, chunk = item - mc * chunks;
        // matrix axes -> matrix id and base offset
        long long rem = mc, mbase = md.base_off;
        int mat = 0;
        for (int a = md.nM - 1; a >= 0; --a) {
            const int c = (int)(rem % md.Mshape[a]);
            rem /= md.Mshape[a];
            mat += c * md.Mmat[a];
            mbase += c * md.Mstride[a];
        }
        __syncthreads();                        // previous item's matrix no longer in use
        const double *msrc = md.mat + (long long)mat * n * n;
        const int n4 = (n + 3) & ~3;            // rows padded to a multiple of 4 (zero rows)
        for (int e = threadIdx.x; e < n4 * NMAX; e += blockDim.x) {
            const int i = e / NMAX, j = e - i * NMAX;
            smat[e] = (i < n && j < n) ? (md.colscale ? msrc[i * n + j] * md.colscale[j] : msrc[i * n + j]) : 0.0;
        }
        __syncthreads();
        const long long f = chunk * blockDim.x + threadIdx.x;
        if (f < md.Fcount) {
            long long r2 = f, base = mbase;
            for (int a = md.nF - 1; a >= 0; --a) {
                const int c = (int)(r2 % md.Fshape[a]);
                r2 /= md.Fshape[a];
                base += c * md.Fstride[a];
            }
            double x[NMAX];
#pragma unroll
            for (int j = 0; j < NMAX; ++j) x[j] = (j < n) ? load(base + j * md.stride) : 0.0;
            // four output rows at a time: eight independent FMA chains hide the fp64 latency
            for (int i = 0; i < n; i += 4) {
                const double2 *r0 = reinterpret_cast<const double2 *>(smat + i * NMAX);
                const double2 *r1 = r0 + NMAX / 2, *r2 = r1 + NMAX / 2, *r3 = r2 + NMAX / 2;
                double a[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                for (int j = 0; j < NMAX / 2; ++j) {
                    const double2 m0 = r0[j], m1 = r1[j], m2 = r2[j], m3 = r3[j];
                    const double xa = x[2 * j], xb = x[2 * j + 1];
                    a[0][0] = fma(m0.x, xa, a[0][0]); a[0][1] = fma(m0.y, xb, a[0][1]);
                    a[1][0] = fma(m1.x, xa, a[1][0]); a[1][1] = fma(m1.y, xb, a[1][1]);
                    a[2][0] = fma(m2.x, xa, a[2][0]); a[2][1] = fma(m2.y, xb, a[2][1]);
                    a[3][0] = fma(m3.x, xa, a[3][0]); a[3][1] = fma(m3.y, xb, a[3][1]);
                }
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (i + r < n) sink(base + (i + r) * md.stride, a[r][0] + a[r][1]);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Tensor-core mode contraction (DMMA m8n8k4) for KRON_TC_MIN <= n <= 64.
// One mode is the GEMM  out[f, i] = sum_k in[f, k] M[i, k]  over the fibres f that share the factor
// matrix M (n x n).  A warp owns a tile of 8 fibres: the A fragments are the fibre values themselves
// (lane l holds in[fibre l/4][k0 + l%4], read straight from global: for the innermost axis that is
// 32 contiguous bytes per fibre, for the other axes 64 contiguous bytes per k, so every sector fetched
// is fully used), the B fragments come from the matrix staged in shared memory with row pitch
// NMAX + 4 (= 4 mod 8 doubles: the 8 x 4 fragment read is bank-conflict free), and the
// ceil(n/8) accumulator tiles stay in registers.  Per fibre tile: ceil(n/4) loads, ceil(n/4) ceil(n/8)
// DMMAs and as many 8-byte LDS, 2 ceil(n/8) stores - against n^2/2 LDS.128 and n^2 DFMA per fibre
// in kron_mode_fibre, whose shared-memory pipe saturates at ~27 % of the fp64 rate.
// Work distribution: the tiles (matrix-axes combination major, fibre tile minor) are split into one
// contiguous range per CTA, so a CTA re-stages the matrix only when its range crosses into the next
// combination, and every CTA gets the same number of tiles (+-1) whatever the grid.  Inside a
// range the warps take tiles round-robin and prefetch the next tile's fragments before the DMMAs
// of the current one.
// ---------------------------------------------------------------------------
// RECT: the output rows are the sub-range [out0, out0 + nout) of the contracted axis (the leading mode of a
// slab-sharded view: this rank forms only its own rows from the full input); IT covers nout, the k steps
// cover n (up to 64, predicated).
//
// Fragment traffic: the A fragments of the NEXT tile are fetched with cp.async (LDGSTS, 8 bytes per lane and
// k step) straight into a per-warp shared-memory stage ([2][16][32] doubles, lane-private columns: no
// synchronisation beyond cp.async.wait_group) while the current tile's DMMAs run, and are read back one LDS
// per k step.  The prefetch therefore costs no registers: the kernel keeps only the 2 IT accumulators live
// (the register-prefetching version needed 4 IT + 2 IT doubles and ran into the 128-register wall of two
// CTAs per SM), and the same stage carries the loader / sink hooks above.
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc, bool valid) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;                            // 0: nothing is read, the 8 bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gsrc), "r"(sz) : "memory");
}
#ifndef KRON_STAGE_BUFS
#define KRON_STAGE_BUFS 2          // fragment stage buffers per warp: 2 = the next tile is fetched while the current one is
#endif                             // contracted; 1 = fetched once the current fragments are consumed (half the shared memory: a third CTA per SM)
#define KRON_STAGE_DOUBLES_PER_WARP (KRON_STAGE_BUFS * 16 * 32)
template <int IT /* 8-row output tiles */, bool EXACT /* n > 8 (IT - 1): straight-line k loop */, bool RECT, class Load, class Sink>
__device__ __forceinline__ void kron_mode_dmma(const KronView &kv, int m, Load &&load, double *smat, Sink &&sink,
                                               KronShare share = KronShare()) {
    constexpr int KT = RECT ? 16 : 2 * IT;                  // n <= 8 IT  =>  ceil(n/4) <= 2 IT
    constexpr int PITCH = 4 * KT + 4;
    const KronMode &md = kv.modes[m];
    const int n = kv.shape[md.dim];
    const int nout = RECT ? md.nout : n, out0 = RECT ? md.out0 : 0;
    const int kt_n = (n + 3) >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    typedef typename kron_bare<Load>::type LoadT;
    typedef typename kron_bare<Sink>::type SinkT;
    constexpr bool XFORM = kron_load_traits<LoadT>::xform, STAGED = kron_sink_traits<SinkT>::staged;
    double *stw = share.stage + warp * KRON_STAGE_DOUBLES_PER_WARP + lane;     // [buf][kt][lane]
    // free-axis decode tables in shared memory: kv.modes[m] with a run-time m is an indexed constant-bank load,
    // and a chain of those per tile (shape, stride per axis) showed up as 20 % of the stall samples of the modes
    __shared__ unsigned s_fshape[SDFS_MAX_DIMS];
    __shared__ long long s_fstride[SDFS_MAX_DIMS];
    // division by the (run-time, per-mode constant) free-axis lengths as multiply-high + shift: for d >= 1 and
    // s = 31 + ceil(log2 d), M = floor(2^s / d) + 1 fits 32 bits and (r M) >> s = floor(r / d) for every r < 2^31
    // (M d - 2^s lies in (0, d], so r (M d - 2^s) < 2^31 d <= 2^s).  ~5 instructions per axis instead of the
    // ~25 of a 32-bit hardware-assisted division; the decode runs once per fibre tile and lane.
    __shared__ unsigned s_fmagic[SDFS_MAX_DIMS];
    __shared__ int s_fsh[SDFS_MAX_DIMS];
    bool tables_written = false;      // written inside the first matrix staging (between its two barriers)
    const int nF = md.nF;
    const long long Fcount = md.Fcount;
    const long long tpm = (md.Fcount + 7) >> 3;              // fibre tiles per matrix combination
    const long long T = md.Mcount * tpm;
    const double *colscale = md.colscale;
    const double *mat0 = md.mat;
    const long long t_begin = T * share.cta / share.nctas, t_end = T * (share.cta + 1) / share.nctas;
    const long long kstride = md.stride;
    const bool small_f = Fcount < (1LL << 31);            // 32-bit index decode (always, in practice)
    const int ktv_lane = (n - q + 3) >> 2;                // number of k steps with 4 kt + q < n for this lane
    int cur_mat = -1;
    for (long long seg = t_begin; seg < t_end;) {
        const long long mc = seg / tpm;
        const long long seg_end = (mc + 1) * tpm < t_end ? (mc + 1) * tpm : t_end;
        long long rem = mc, mbase = md.base_off;
        int mat = 0;
        for (int a = md.nM - 1; a >= 0; --a) {
            const int c = (int)(rem % md.Mshape[a]);
            rem /= md.Mshape[a];
            mat += c * md.Mmat[a];
            mbase += c * md.Mstride[a];
        }
        if (mat != cur_mat) {                   // uniform over the CTA
            __syncthreads();                    // previous matrix (and a previous call's decode tables) no longer in use
            if (!tables_written) {
                if (threadIdx.x < SDFS_MAX_DIMS) {
                    const unsigned d = threadIdx.x < nF ? (unsigned)md.Fshape[threadIdx.x] : 1u;
                    const int sh = 31 + (d > 1 ? 32 - __clz((int)(d - 1)) : 0);
                    s_fshape[threadIdx.x] = d;
                    s_fstride[threadIdx.x] = threadIdx.x < nF ? md.Fstride[threadIdx.x] : 0;
                    s_fmagic[threadIdx.x] = (unsigned)((1ULL << sh) / d + 1ULL);
                    s_fsh[threadIdx.x] = sh;
                }
                tables_written = true;
            }
            const double *msrc = mat0 + (long long)mat * n * n + (long long)out0 * n;
            for (int e0 = threadIdx.x; e0 < IT * 8 * PITCH; e0 += 4 * blockDim.x) {      // four loads in flight per thread
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * blockDim.x;
                    const int i = e / PITCH, j = e - i * PITCH;
                    v[u] = (e < IT * 8 * PITCH && i < nout && j < n) ? msrc[i * n + j] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * blockDim.x;
                    const int j = e % PITCH;
                    if (e < IT * 8 * PITCH) smat[e] = (colscale && j < n) ? v[u] * colscale[j] : v[u];
                }
            }
            __syncthreads();
            cur_mat = mat;
        }
        // fragment fetch: lane (g, q) copies fibre 8 t + g at k = 4 kt + q into stage buffer `buf`
        auto fetch_tile = [&](long long t, int buf, long long &base, bool &fv) {
            const long long f = (t - mc * tpm) * 8 + g;
            fv = f < Fcount;
            base = mbase;
            if (small_f) {
                unsigned r2 = fv ? (unsigned)f : 0u;
                for (int ax = nF - 1; ax >= 0; --ax) {
                    const unsigned qd = (unsigned)(((unsigned long long)r2 * s_fmagic[ax]) >> s_fsh[ax]);
                    base += (long long)(r2 - qd * s_fshape[ax]) * s_fstride[ax];
                    r2 = qd;
                }
            } else {
                long long r2 = fv ? f : 0;
                for (int ax = nF - 1; ax >= 0; --ax) {
                    const int c = (int)(r2 % s_fshape[ax]);
                    r2 /= s_fshape[ax];
                    base += c * s_fstride[ax];
                }
            }
            const double *p = load.ptr(base + q * kstride);
            double *dst = stw + buf * (16 * 32);
            const int ktv = fv ? ktv_lane : 0;       // k steps this lane has data for (4 kt + q < n), none for a padding fibre
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                if (EXACT && !RECT ? (kt < KT - 1 || kt < kt_n) : kt < kt_n)
                    cp_async8(dst + kt * 32, p, kt < ktv);
                p += 4 * kstride;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        long long t = seg + warp;
        long long base = 0, bn = 0;
        bool fv = false, vn = false;
        int buf = 0;
        if (t < seg_end) fetch_tile(t, 0, base, fv);
        const double *brow = smat + g * PITCH + q;
#ifdef KRON_DBG_NO_B_LDS
        const double dbg_b = brow[0];
#endif
        while (t < seg_end) {
            const long long tn = t + nwarps;
            if (KRON_STAGE_BUFS == 2 && tn < seg_end) {
                fetch_tile(tn, buf ^ 1, bn, vn);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            double *st = stw + buf * (16 * 32);
            if constexpr (XFORM) {              // loader arithmetic in place, rolled: four independent chains per trip
                if (fv && load.active()) {
#pragma unroll 1
                    for (int kt = 0; kt < kt_n; kt += 4) {
                        double x[4];
                        bool v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            v[u] = kt + u < kt_n && 4 * (kt + u) + q < n;
                            x[u] = v[u] ? st[(kt + u) * 32] : 1.0;
                        }
                        const long long p0 = base + (long long)(4 * kt + q) * kstride;
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = load.xform(v[u] ? p0 + 4LL * u * kstride : p0, x[u]);
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (v[u]) st[(kt + u) * 32] = x[u];
                    }
                }
            }
            double c[IT][2];
#pragma unroll
            for (int it = 0; it < IT; ++it) c[it][0] = c[it][1] = 0.0;
            // n > 8 (IT - 1) for the exact instantiations, so only the last k step can be absent: the others
            // form one straight-line block; the even-count and restricted-output variants predicate every step
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                if (EXACT && !RECT ? (kt < KT - 1 || kt < kt_n) : kt < kt_n) {
                    const double a = st[kt * 32];
#ifdef KRON_DBG_NO_B_LDS      /* timing experiment only (wrong results): B operands from registers, no LDS in the DMMA loop */
#pragma unroll
                    for (int it = 0; it < IT; ++it) dmma884(c[it][0], c[it][1], a, (it & 1) ? a : dbg_b);
#else
#pragma unroll
                    for (int it = 0; it < IT; ++it) dmma884(c[it][0], c[it][1], a, brow[it * 8 * PITCH + kt * 4]);
#endif
                }
            }
            if (KRON_STAGE_BUFS == 1 && !STAGED && tn < seg_end) fetch_tile(tn, 0, bn, vn);     // fragments consumed: refill under the stores
            if constexpr (STAGED) {               // sink arithmetic from the stage, rolled (four outputs per trip)
                // the sink's per-row factor (a_row) for all 2 IT outputs first: every load in flight at once
                // instead of one exposed latency per trip; c is dead after this block
                if (fv) {
                    long long idx = base + (out0 + 2 * q) * kstride;
                    double f[IT][2];
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int i = it * 8 + 2 * q;
                        f[it][0] = i < nout ? sink.pre(idx) : 1.0;
                        f[it][1] = i + 1 < nout ? sink.pre(idx + kstride) : 1.0;
                        idx += 8 * kstride;
                    }
#pragma unroll
                    for (int it = 0; it < IT; ++it) { st[(2 * it) * 32] = c[it][0] * f[it][0]; st[(2 * it + 1) * 32] = c[it][1] * f[it][1]; }
                }
                // innermost-axis mode with a single-output sink: the values go back into the stage and the warp
                // stores each fibre's nout contiguous outputs with 16-byte stores (448 contiguous bytes per fibre
                // at n = 56 instead of 8-byte stores in 64-byte pieces: what NVLink peer stores and HBM both want)
                const bool wide = kstride == 1 && sink.single_output();
                if (fv) {
                    long long idx = base + (out0 + 2 * q) * kstride;
                    const int it_o = (nout + 7) >> 3;
#pragma unroll 1
                    for (int it = 0; it < it_o; it += 2) {
                        const int i = it * 8 + 2 * q;
                        double s0 = st[(2 * it) * 32], s1 = st[(2 * it + 1) * 32];
                        double s2 = st[(2 * it + 2) * 32], s3 = st[(2 * it + 3) * 32];
                        // rows i, i+1, i+8, i+9 (the last ones may not exist)
                        if (!sink.single_output()) {
                            sink.quad(idx, kstride, s0, s1, s2, s3, nout - i);
                        } else {
                            sink.values(idx, kstride, s0, s1, s2, s3, nout - i);
                            if (wide) {
                                st[(2 * it) * 32] = s0; st[(2 * it + 1) * 32] = s1;
                                st[(2 * it + 2) * 32] = s2; st[(2 * it + 3) * 32] = s3;
                            } else {
                                const int left = nout - i;
                                if (left > 0) sink.store1(idx, s0);
                                if (left > 1) sink.store1(idx + kstride, s1);
                                if (left > 8) sink.store1(idx + 8 * kstride, s2);
                                if (left > 9) sink.store1(idx + 9 * kstride, s3);
                            }
                        }
                        idx += 16 * kstride;
                    }
                }
                if (wide) {
                    __syncwarp();
                    const double *sw = st - lane;                     // this warp's stage buffer, all lanes' columns
#pragma unroll 1
                    for (int f8 = 0; f8 < 8; ++f8) {
                        const long long bf = __shfl_sync(0xffffffffu, base, 4 * f8) + out0;
                        const int vf = __shfl_sync(0xffffffffu, (int)fv, 4 * f8);
                        const int i0 = 2 * lane;
                        if (vf && i0 < nout) {
                            const int it = i0 >> 3, qq = (i0 & 7) >> 1;
                            const double v0 = sw[(2 * it) * 32 + 4 * f8 + qq], v1 = sw[(2 * it + 1) * 32 + 4 * f8 + qq];
                            if (i0 + 1 < nout && ((bf & 1) == 0)) sink.store2(bf + i0, v0, v1);
                            else { sink.store1(bf + i0, v0); if (i0 + 1 < nout) sink.store1(bf + i0 + 1, v1); }
                        }
                    }
                    __syncwarp();
                }
            } else if (fv) {
                long long idx = base + (out0 + 2 * q) * kstride;
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    const int i = it * 8 + 2 * q;
                    if (i < nout) sink(idx, c[it][0]);
                    if (i + 1 < nout) sink(idx + kstride, c[it][1]);
                    idx += 8 * kstride;
                }
            }
            if (KRON_STAGE_BUFS == 1 && STAGED && tn < seg_end) fetch_tile(tn, 0, bn, vn);      // the staged sink used the buffer
            base = bn; fv = vn;
            if (KRON_STAGE_BUFS == 2) buf ^= 1;
            t = tn;
        }
        seg = seg_end;
    }
}

// Register-fed variant of the tensor-core contraction (fragments loaded with plain loads, no stage): for
// vectors that already live in SHARED memory (the fused sweep kernel contracts a column in place), where
// cp.async has no global source to copy from and latency is not an issue.
template <int IT /* 8-row output tiles */, bool PREFETCH, class Sink>
__device__ __forceinline__ void kron_mode_dmma_reg(const KronView &kv, int m, const double *in, double *smat, Sink &&sink,
                                               KronShare share = KronShare()) {
    constexpr int PITCH = 8 * IT + 4, KT = 2 * IT;          // n <= 8 IT  =>  ceil(n/4) <= 2 IT
    const KronMode &md = kv.modes[m];
    const int n = kv.shape[md.dim];
    const int kt_n = (n + 3) >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const long long tpm = (md.Fcount + 7) >> 3;              // fibre tiles per matrix combination
    const long long T = md.Mcount * tpm;
    const long long t_begin = T * share.cta / share.nctas, t_end = T * (share.cta + 1) / share.nctas;
    const long long kstride = md.stride;
    const bool small_f = md.Fcount < (1LL << 31);            // 32-bit index decode (always, in practice)
    int cur_mat = -1;
    for (long long seg = t_begin; seg < t_end;) {
        const long long mc = seg / tpm;
        const long long seg_end = (mc + 1) * tpm < t_end ? (mc + 1) * tpm : t_end;
        long long rem = mc, mbase = md.base_off;
        int mat = 0;
        for (int a = md.nM - 1; a >= 0; --a) {
            const int c = (int)(rem % md.Mshape[a]);
            rem /= md.Mshape[a];
            mat += c * md.Mmat[a];
            mbase += c * md.Mstride[a];
        }
        if (mat != cur_mat) {                   // uniform over the CTA
            __syncthreads();                    // previous matrix no longer in use
            const double *msrc = md.mat + (long long)mat * n * n;
            for (int e = threadIdx.x; e < IT * 8 * PITCH; e += blockDim.x) {
                const int i = e / PITCH, j = e - i * PITCH;
                smat[e] = (i < n && j < n) ? msrc[i * n + j] : 0.0;
            }
            __syncthreads();
            cur_mat = mat;
        }
        // fragment loader: lane (g, q) reads fibre 8 t + g at k = 4 kt + q
        auto load_tile = [&](long long t, double (&a)[KT], long long &base, bool &fv) {
            const long long f = (t - mc * tpm) * 8 + g;
            fv = f < md.Fcount;
            base = mbase;
            if (small_f) {
                unsigned r2 = fv ? (unsigned)f : 0u;
                for (int ax = md.nF - 1; ax >= 0; --ax) {
                    const unsigned sh = (unsigned)md.Fshape[ax], qd = r2 / sh;
                    base += (long long)(r2 - qd * sh) * md.Fstride[ax];
                    r2 = qd;
                }
            } else {
                long long r2 = fv ? f : 0;
                for (int ax = md.nF - 1; ax >= 0; --ax) {
                    const int c = (int)(r2 % md.Fshape[ax]);
                    r2 /= md.Fshape[ax];
                    base += c * md.Fstride[ax];
                }
            }
            const double *p = in + base + q * kstride;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                a[kt] = (fv && kt * 4 + q < n) ? *p : 0.0;
                p += 4 * kstride;
            }
        };
        long long t = seg + warp;
        double a[KT];
        long long base = 0;
        bool fv = false;
        if (PREFETCH && t < seg_end) load_tile(t, a, base, fv);
        const double *brow = smat + g * PITCH + q;
        while (t < seg_end) {
            double an[KT];
            long long bn = 0;
            bool vn = false;
            const long long tn = t + nwarps;
            if (!PREFETCH) load_tile(t, a, base, fv);
            if (PREFETCH && tn < seg_end) load_tile(tn, an, bn, vn);
            else if (PREFETCH) {
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) an[kt] = 0.0;
            }
            double c[IT][2];
#pragma unroll
            for (int it = 0; it < IT; ++it) c[it][0] = c[it][1] = 0.0;
            // n > 8 (IT - 1) for the exact instantiations (PREFETCH) and n > 8 (IT - 2) for the even ones, so
            // only the last 1 (3) k steps can be absent: the others form one straight-line block
            constexpr int KT_SURE = PREFETCH ? KT - 1 : KT - 3;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                if (kt < KT_SURE || kt < kt_n) {
#pragma unroll
                    for (int it = 0; it < IT; ++it) dmma884(c[it][0], c[it][1], a[kt], brow[it * 8 * PITCH + kt * 4]);
                }
            }
            if (fv) {
                long long idx = base + 2 * q * kstride;
#pragma unroll
                for (int it = 0; it < IT; ++it) {
                    const int i = it * 8 + 2 * q;
                    if (i < n) sink(idx, c[it][0]);
                    if (i + 1 < n) sink(idx + kstride, c[it][1]);
                    idx += 8 * kstride;
                }
            }
            if (PREFETCH) {
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) a[kt] = an[kt];
                base = bn; fv = vn;
            }
            t = tn;
        }
        seg = seg_end;
    }
}

// Short axes (9 <= n <= 16, two output tiles): a tile is only ~6 DMMAs, so the loop is bound by the
// latency of its 3-4 fragment loads.  TP tiles per warp iteration, all their loads issued before the
// first DMMA, keep TP x more bytes in flight (batched sweep panels, N = 10^4 x 4096 columns:
// 0.31 -> 0.1x ms per mode).
template <int IT, int TP, class Load, class Sink>
__device__ __forceinline__ void kron_mode_dmma_multi(const KronView &kv, int m, Load &&load, double *smat, Sink &&sink,
                                                     KronShare share = KronShare()) {
    constexpr int PITCH = 8 * IT + 4, KT = 2 * IT;
    const KronMode &md = kv.modes[m];
    const int n = kv.shape[md.dim];
    const int kt_n = (n + 3) >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const long long tpm = (md.Fcount + 7) >> 3;
    const long long T = md.Mcount * tpm;
    const long long t_begin = T * share.cta / share.nctas, t_end = T * (share.cta + 1) / share.nctas;
    const long long kstride = md.stride;
    const bool small_f = md.Fcount < (1LL << 31);
    int cur_mat = -1;
    for (long long seg = t_begin; seg < t_end;) {
        const long long mc = seg / tpm;
        const long long seg_end = (mc + 1) * tpm < t_end ? (mc + 1) * tpm : t_end;
        long long rem = mc, mbase = md.base_off;
        int mat = 0;
        for (int a = md.nM - 1; a >= 0; --a) {
            const int c = (int)(rem % md.Mshape[a]);
            rem /= md.Mshape[a];
            mat += c * md.Mmat[a];
            mbase += c * md.Mstride[a];
        }
        if (mat != cur_mat) {
            __syncthreads();
            const double *msrc = md.mat + (long long)mat * n * n;
            for (int e = threadIdx.x; e < IT * 8 * PITCH; e += blockDim.x) {
                const int i = e / PITCH, j = e - i * PITCH;
                smat[e] = (i < n && j < n) ? (md.colscale ? msrc[i * n + j] * md.colscale[j] : msrc[i * n + j]) : 0.0;
            }
            __syncthreads();
            cur_mat = mat;
        }
        const double *brow = smat + g * PITCH + q;
        for (long long t = seg + warp; t < seg_end; t += (long long)nwarps * TP) {
            double a[TP][KT];
            long long base[TP];
            bool fv[TP];
#pragma unroll
            for (int u = 0; u < TP; ++u) {
                const long long tt = t + (long long)u * nwarps;
                const long long f = (tt - mc * tpm) * 8 + g;
                fv[u] = tt < seg_end && f < md.Fcount;
                base[u] = mbase;
                if (small_f) {
                    unsigned r2 = fv[u] ? (unsigned)f : 0u;
                    for (int ax = md.nF - 1; ax >= 0; --ax) {
                        const unsigned sh = (unsigned)md.Fshape[ax], qd = r2 / sh;
                        base[u] += (long long)(r2 - qd * sh) * md.Fstride[ax];
                        r2 = qd;
                    }
                } else {
                    long long r2 = fv[u] ? f : 0;
                    for (int ax = md.nF - 1; ax >= 0; --ax) {
                        const int c = (int)(r2 % md.Fshape[ax]);
                        r2 /= md.Fshape[ax];
                        base[u] += c * md.Fstride[ax];
                    }
                }
                long long p = base[u] + q * kstride;
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    a[u][kt] = (fv[u] && kt * 4 + q < n) ? load(p) : 0.0;
                    p += 4 * kstride;
                }
            }
#pragma unroll
            for (int u = 0; u < TP; ++u) {
                double c[IT][2];
#pragma unroll
                for (int it = 0; it < IT; ++it) c[it][0] = c[it][1] = 0.0;
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    if (kt < KT - 1 || kt < kt_n) {          // n >= 9: at least KT - 1 = 3 k steps
#pragma unroll
                        for (int it = 0; it < IT; ++it) dmma884(c[it][0], c[it][1], a[u][kt], brow[it * 8 * PITCH + kt * 4]);
                    }
                }
                if (fv[u]) {
                    long long idx = base[u] + 2 * q * kstride;
#pragma unroll
                    for (int it = 0; it < IT; ++it) {
                        const int i = it * 8 + 2 * q;
                        if (i < n) sink(idx, c[it][0]);
                        if (i + 1 < n) sink(idx + kstride, c[it][1]);
                        idx += 8 * kstride;
                    }
                }
            }
        }
        seg = seg_end;
    }
}

// dispatch on the size of the contracted axis: register-tiled FMA kernel for short axes, the
// tensor-core contraction up to 64, the cached-load pass beyond
// (PREFETCH: software-pipelined fragment loads, +32 registers - for the stand-alone mode kernels;
// the persistent loop kernels keep the lean variant)
// (SMALL_FMA: thread-per-fibre FMA contraction up to n = 16 - for vectors resident in shared memory,
// where coalescing is moot and a warp-wide tile of only 8 fibres costs ~5x the instructions per fibre)
// (RECT_OK: also instantiate the restricted-output variants used by the leading mode of a slab-sharded view;
// only call sites that may contract such a mode pay for the extra code)
template <bool PREFETCH = false, bool SMALL_FMA = false, bool RECT_OK = false, class Load, class Sink>
__device__ __forceinline__ void kron_mode_apply_ld(const KronView &kv, int m, Load &&load, double *smat, Sink &&sink,
                                                   KronShare share = KronShare()) {
    const KronMode &md = kv.modes[m];
    const int n = kv.shape[md.dim];
    constexpr bool HOOKS = kron_load_traits<typename kron_bare<Load>::type>::xform ||
                           kron_sink_traits<typename kron_bare<Sink>::type>::staged;
    if (md.nout != n) {                         // restricted output rows (validated on the host: 9 <= n <= 64)
        if constexpr (RECT_OK) {
            const int it_o = (md.nout + 7) >> 3;
            if (it_o <= 1) kron_mode_dmma<1, PREFETCH, true>(kv, m, load, smat, sink, share);
            else if (it_o <= 2) kron_mode_dmma<2, PREFETCH, true>(kv, m, load, smat, sink, share);
            else if (it_o <= 4) kron_mode_dmma<4, PREFETCH, true>(kv, m, load, smat, sink, share);
            else kron_mode_dmma<8, PREFETCH, true>(kv, m, load, smat, sink, share);
        }
        return;
    }
    if constexpr (SMALL_FMA) {                  // vector resident in shared memory: no cp.async staging
        if (n > 8 && n <= 16) { kron_mode_fibre<16>(kv, m, load, smat, sink, share); return; }
        if (n > 16 && n <= KRON_NMAX_LIMIT) {
            const int it_s = (n + 7) >> 3;
            if (it_s <= 4) kron_mode_dmma_reg<4, false>(kv, m, load.ptr(0), smat, sink, share);
            else if (it_s <= 6) kron_mode_dmma_reg<6, false>(kv, m, load.ptr(0), smat, sink, share);
            else kron_mode_dmma_reg<8, false>(kv, m, load.ptr(0), smat, sink, share);
            return;
        }
    }
    if (n < KRON_TC_MIN) kron_mode_fibre<8>(kv, m, load, smat, sink, share);      // n <= 8
    else if (n <= KRON_NMAX_LIMIT) {
        const int it_n = (n + 7) >> 3;          // 2..8 output tiles
        if constexpr (PREFETCH) {               // stand-alone kernels: exact tile count
            switch (it_n) {
            case 2:
                if constexpr (HOOKS) kron_mode_dmma<2, true, false>(kv, m, load, smat, sink, share);
                else kron_mode_dmma_multi<2, 4>(kv, m, load, smat, sink, share);
                break;
            case 3: kron_mode_dmma<3, true, false>(kv, m, load, smat, sink, share); break;
            case 4: kron_mode_dmma<4, true, false>(kv, m, load, smat, sink, share); break;
            case 5: kron_mode_dmma<5, true, false>(kv, m, load, smat, sink, share); break;
            case 6: kron_mode_dmma<6, true, false>(kv, m, load, smat, sink, share); break;
            case 7: kron_mode_dmma<7, true, false>(kv, m, load, smat, sink, share); break;
            default: kron_mode_dmma<8, true, false>(kv, m, load, smat, sink, share); break;
            }
        } else {                                // loop kernels: even tile counts (zero-padded rows)
            if (it_n <= 2) {
                if constexpr (HOOKS) kron_mode_dmma<2, false, false>(kv, m, load, smat, sink, share);
                else kron_mode_dmma_multi<2, 4>(kv, m, load, smat, sink, share);
            }
            else if (it_n <= 4) kron_mode_dmma<4, false, false>(kv, m, load, smat, sink, share);
            else if (it_n <= 6) kron_mode_dmma<6, false, false>(kv, m, load, smat, sink, share);
            else kron_mode_dmma<8, false, false>(kv, m, load, smat, sink, share);
        }
    }
    else kron_mode_pass(kv, m, load, (int64_t)share.cta * blockDim.x + threadIdx.x, (int64_t)share.nctas * blockDim.x, sink);
}
template <bool PREFETCH = false, bool SMALL_FMA = false, bool RECT_OK = false, class Sink>
__device__ __forceinline__ void kron_mode_apply(const KronView &kv, int m, const double *in, double *smat, Sink &&sink,
                                                KronShare share = KronShare()) {
    kron_mode_apply_ld<PREFETCH, SMALL_FMA, RECT_OK>(kv, m, KronLoadPlain{in}, smat, sink, share);
}
// Stand-alone mode kernel (k_kron_mode: one launch per mode - the large-operator path and the batched sweep):
// a launch has the whole register file for two 256-thread CTAs per SM, so the register-prefetching variant
// (next tile's fragments held in registers under the current tile's DMMAs) fits, and it is ~10 % faster there
// than the cp.async-staged one (54 against 62 us per mode at 9.8 M states); the staged variant is what the
// persistent kernels use, where registers are the scarce resource.
template <class Sink>
__device__ __forceinline__ void kron_mode_apply_regpf(const KronView &kv, int m, const double *in, double *smat, Sink &&sink,
                                                      KronShare share) {
    const KronMode &md = kv.modes[m];
    const int n = kv.shape[md.dim];
    if (md.nout != n || md.colscale || n < KRON_TC_MIN || n > KRON_NMAX_LIMIT || ((n + 7) >> 3) == 2) {
        kron_mode_apply<true, false, true>(kv, m, in, smat, sink, share);
        return;
    }
    switch ((n + 7) >> 3) {
    case 3: kron_mode_dmma_reg<3, true>(kv, m, in, smat, sink, share); break;
    case 4: kron_mode_dmma_reg<4, true>(kv, m, in, smat, sink, share); break;
    case 5: kron_mode_dmma_reg<5, true>(kv, m, in, smat, sink, share); break;
    case 6: kron_mode_dmma_reg<6, true>(kv, m, in, smat, sink, share); break;
    case 7: kron_mode_dmma_reg<7, true>(kv, m, in, smat, sink, share); break;
    default: kron_mode_dmma_reg<8, true>(kv, m, in, smat, sink, share); break;
    }
}
#define KRON_SMAT_DOUBLES (KRON_NMAX_LIMIT * (KRON_NMAX_LIMIT + 4))
static_assert(KRON_NMAX_LIMIT == 64, "the restricted-output contraction stages 8 IT x (4 x 16 + 4) doubles");

// Out-of-line storing contraction for the persistent loop kernels: they apply the operator at many
// sites (T, JVP inside BiCGSTAB / GMRES, Anderson), and one shared copy of the non-final modes keeps
// their code size (instruction-cache footprint, compile time) bounded.
static __device__ __noinline__ void kron_mode_store(const KronView &kv, int m, const double *in, double *out, double *smat,
                                                    double *stage) {
    kron_mode_apply<true, false, true>(kv, m, in, smat, [&](int64_t idx, double s) { out[idx] = s; }, KronShare(stage));   // exact tile counts: one copy
}
