// Streaming row-tile dot products: the HBM-bound core of every dense operator
// application  (y = P x with fused prologue/epilogue).
//
// Data layout: P row-major fp64, leading dimension ld (rows 16-byte aligned when
// ld is even -> 16-byte vector loads; otherwise 8-byte loads).  One warp owns R
// consecutive rows and walks the columns in 64-column chunks: lane l reads columns
// 64c+2l, 64c+2l+1 of every row as one 16-byte streaming load (512 contiguous
// bytes per row per warp instruction), U chunks are issued back to back so each
// lane keeps R*U independent 16-byte loads in flight.  x comes through L1/L2 with
// default caching and is shared by the R rows.  Per-lane partial sums are
// accumulated in a fixed column order and reduced with a fixed xor-shuffle tree,
// so a row's result is bit-identical for every R, U, grid size and row sharding.
#pragma once
#include "common.cuh"

template <int R, int U, int NX, bool VEC2>
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
                if (VEC2) pv[u][r] = ld_stream2(p);
                else pv[u][r] = make_double2(ld_stream1(p), ld_stream1(p + 1));
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xv[0][u] = ld_x2(xa + ((c + u) << 6));
            if (NX > 1) xv[NX - 1][u] = ld_x2(xb + ((c + u) << 6));
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
        double2 pv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double *p = pr[r] + (c << 6);
            if (VEC2) pv[r] = ld_stream2(p);
            else pv[r] = make_double2(ld_stream1(p), ld_stream1(p + 1));
        }
        double2 xv[NX];
        xv[0] = ld_x2(xa + (c << 6));
        if (NX > 1) xv[NX - 1] = ld_x2(xb + (c << 6));
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int q = 0; q < NX; ++q) {
                a[q][r][0] = fma(pv[r].x, xv[q].x, a[q][r][0]);
                a[q][r][1] = fma(pv[r].y, xv[q].y, a[q][r][1]);
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

// Select the value belonging to lane `lane` (< R) out of a register array.
template <int R>
__device__ __forceinline__ double pick(const double (&v)[R], int lane) {
    double s = v[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
        if (lane == r) s = v[r];
    return s;
}

// One pass over this rank's rows.  Warp `wg` of `nw` owns the contiguous row range
// [nloc*wg/nw, nloc*(wg+1)/nw); groups of 4 rows, then 2, then 1, so no lane ever
// touches an invalid row.  epi(n_global, s0, s1) runs on one lane per row.
template <int NX, class Epi>
__device__ __forceinline__ void dense_rows_pass(const DenseView &dv, const double *x0,
                                                const double *x1, int wg, int nw,
                                                Epi &&epi) {
    const int lane = threadIdx.x & 31;
    const int64_t nloc = dv.row_end - dv.row_begin;
    int64_t r = nloc * wg / nw;
    const int64_t r1 = nloc * (wg + 1) / nw;
    constexpr int R4U = (NX == 1) ? 4 : 2;
    for (; r + 4 <= r1; r += 4) {
        double s[NX][4];
        if (dv.vec2) warp_rows_dot<4, R4U, NX, true>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
        else warp_rows_dot<4, R4U, NX, false>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
        if (lane < 4) epi(dv.row_begin + r + lane, pick<4>(s[0], lane), pick<4>(s[NX - 1], lane));
    }
    if (r + 2 <= r1) {
        double s[NX][2];
        if (dv.vec2) warp_rows_dot<2, 4, NX, true>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
        else warp_rows_dot<2, 4, NX, false>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
        if (lane < 2) epi(dv.row_begin + r + lane, pick<2>(s[0], lane), pick<2>(s[NX - 1], lane));
        r += 2;
    }
    if (r < r1) {
        double s[NX][1];
        if (dv.vec2) warp_rows_dot<1, 8, NX, true>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
        else warp_rows_dot<1, 8, NX, false>(dv.P + r * dv.ld, dv.ld, x0, x1, dv.N, s);
        if (lane < 1) epi(dv.row_begin + r, s[0][0], s[NX - 1][0]);
    }
}

// Factor-structured apply: one mode contraction
//   out[idx] = sum_j M[mat_id(idx)][i_m][j] * in[idx with i_m := j]
// over a grid-stride range of output elements.  Called once per mode with a grid
// barrier between calls; the last mode feeds the epilogue instead of storing.
template <class Sink>
__device__ __forceinline__ void kron_mode_pass(const KronView &kv, int m,
                                               const double *__restrict__ in, int64_t tid,
                                               int64_t nthreads, Sink &&sink) {
    const KronMode &md = kv.modes[m];
    const int dim = md.dim;
    const int n = kv.shape[dim];
    int64_t stride = 1;
    for (int d = kv.D - 1; d > dim; --d) stride *= kv.shape[d];
    for (int64_t idx = tid; idx < kv.N; idx += nthreads) {
        int64_t rem = idx;
        int mat = 0, im = 0;
        for (int d = kv.D - 1; d >= 0; --d) {
            const int cd = (int)(rem % kv.shape[d]);
            rem /= kv.shape[d];
            mat += cd * md.mstride[d];
            if (d == dim) im = cd;
        }
        const double *row = md.mat + ((int64_t)mat * n + im) * n;
        const double *src = in + (idx - (int64_t)im * stride);
        double acc = 0.0;
        for (int j = 0; j < n; ++j) acc = fma(row[j], src[(int64_t)j * stride], acc);
        sink(idx, acc);
    }
}
