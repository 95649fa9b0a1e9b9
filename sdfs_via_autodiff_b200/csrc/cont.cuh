// Continuous-state operator pass: for every grid state x,
//   Kg(x) = const(x) * sum_q W_q exp(theta h_lam'(x, eta_q)) * interp(w)(x'(x, eta_q))^theta
// (Kg_vmap_quad / Kg_vmap_mc, ssy_wc_ratio_continuous.py:90-153), and its linearisation
//   L(v)(x) = sum_q W_q exp(theta h_lam') interp(w)(x')^(theta-1) interp(v)(x').
// interp = multilinear interpolation on uniform grids with nearest-edge extension
// (utils.py:6-23: jax.scipy.ndimage.map_coordinates(order=1, mode='nearest')).
// One warp per state; lanes stride over the shock nodes; fixed-order warp reduction, so the
// result of a state does not depend on the launch configuration.  Compute/gather bound
// (Q pow + exp and 2^D gathers per state), not an HBM stream.
#pragma once
#include "common.cuh"

template <int D>
__device__ __forceinline__ void cont_decode(const ContView &cv, int64_t n, double (&x)[D]) {
    int64_t rem = n;
#pragma unroll
    for (int d = D - 1; d >= 0; --d) {
        const int i = (int)(rem % cv.n[d]);
        rem /= cv.n[d];
        x[d] = cv.grid[d][i];
    }
}

// const(x) = exp((1-gamma)(mu_c + z) + (1-gamma)^2 sigma_c^2 / 2), sigma_c = phi_c exp(h_c)
template <int D>
__device__ __forceinline__ double cont_const(const ContView &cv, const double (&x)[D]) {
    const double z = (D == 4) ? x[3] : x[4];
    const double sc = cv.phi_c * exp(x[1]);
    const double omg = 1.0 - cv.gamma;
    return exp(omg * (cv.mu_c + z) + 0.5 * omg * omg * sc * sc);
}

// state-dependent volatilities needed by next_state (one exp per state, not per node)
template <int D>
struct ContVol { double s_a, s_b; };
template <int D>
__device__ __forceinline__ ContVol<D> cont_vol(const ContView &cv, const double (&x)[D]) {
    ContVol<D> v;
    if (D == 4) { v.s_a = cv.p[5] * exp(x[2]); v.s_b = 0.0; }                 // sigma_z = phi_z exp(h_z)
    else { v.s_a = cv.p[9] * exp(x[2]); v.s_b = cv.p[15] * exp(x[3]); }       // sigma_z, sigma_zpi
    return v;
}

template <int D>
__device__ __forceinline__ void cont_next(const ContView &cv, const double (&x)[D], const ContVol<D> &vol, int q,
                                          double (&xn)[D]) {
    const double *eta = cv.nodes + q;
    const int Q = cv.Q;
    const double *p = cv.p;
    if (D == 4) {   // SSY params: beta,gamma,psi,mu_c,rho,phi_z,phi_c,rho_z,rho_c,rho_lam,s_z,s_c,s_lam
        xn[0] = p[9] * x[0] + p[12] * eta[0];
        xn[1] = p[8] * x[1] + p[11] * eta[Q];
        xn[2] = p[7] * x[2] + p[10] * eta[2 * Q];
        xn[3] = p[4] * x[3] + vol.s_a * eta[3 * Q];
    } else {        // GCY params: beta,psi,gamma,rho_lam,s_lam,mu_c,phi_c,rho,rho_pi,phi_z,rho_c,s_c,rho_z,s_z,rho_pipi,phi_zpi,rho_zpi,s_zpi
        xn[0] = p[3] * x[0] + p[4] * eta[0];
        xn[1] = p[10] * x[1] + p[11] * eta[Q];
        xn[2] = p[12] * x[2] + p[13] * eta[2 * Q];
        xn[3] = p[16] * x[3] + p[17] * eta[3 * Q];
        xn[4] = p[7] * x[4] + p[8] * x[5] + vol.s_a * eta[4 * Q];
        xn[5] = p[14] * x[5] + vol.s_b * eta[5 * Q];
    }
}

// multilinear interpolation of one or two vectors at the same point
template <int D, int NV>
__device__ __forceinline__ void cont_interp(const ContView &cv, const double (&xn)[D], const double *f0, const double *f1,
                                            double &o0, double &o1) {
    int lo[D], hi[D];
    double t[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        const double c = (xn[d] - cv.g0[d]) / cv.intv[d];
        const double fl = floor(c);
        t[d] = c - fl;
        int i = (int)fl;
        const int m = cv.n[d] - 1;
        lo[d] = i < 0 ? 0 : (i > m ? m : i);
        ++i;
        hi[d] = i < 0 ? 0 : (i > m ? m : i);
    }
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int corner = 0; corner < (1 << D); ++corner) {
        double wgt = 1.0;
        int64_t idx = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const bool up = (corner >> (D - 1 - d)) & 1;
            wgt *= up ? t[d] : (1.0 - t[d]);
            idx = idx * cv.n[d] + (up ? hi[d] : lo[d]);
        }
        a0 = fma(wgt, f0[idx], a0);
        if (NV > 1) a1 = fma(wgt, f1[idx], a1);
    }
    o0 = a0;
    o1 = a1;
}

// mode 0: epi(n, Kg)                         (T pass; wv = w)
// mode 1: epi(n, L(v))                       (linearised pass; wv = w, vv = v)
// mode 2: epi2(n, Kg, L(v))                  (both in one sweep over the nodes)
template <int D, int MODE, class Epi>
__device__ __forceinline__ void cont_pass_d(const ContView &cv, const double *wv, const double *vv, Epi &&epi) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double theta = cv.theta;
    for (int64_t n = cv.row_begin + warp_g; n < cv.row_end; n += nwarps) {
        double x[D], xn[D];
        cont_decode<D>(cv, n, x);
        const ContVol<D> vol = cont_vol<D>(cv, x);
        double accT = 0.0, accJ = 0.0;
        for (int q = lane; q < cv.Q; q += 32) {
            cont_next<D>(cv, x, vol, q, xn);
            const double pf = cv.weights[q] * exp(theta * xn[0]);
            double wi, vi;
            cont_interp<D, (MODE == 0 ? 1 : 2)>(cv, xn, wv, MODE == 0 ? wv : vv, wi, vi);
            if (MODE != 1) accT = fma(pf, pow(wi, theta), accT);
            if (MODE != 0) accJ = fma(pf * pow(wi, theta - 1.0), vi, accJ);
        }
        accT = warp_sum(accT);
        accJ = warp_sum(accJ);
        if (lane == 0) {
            const double cst = (MODE == 1) ? 1.0 : cont_const<D>(cv, x);
            epi(n, cst * accT, accJ);
        }
    }
}

template <int MODE, class Epi>
__device__ __forceinline__ void cont_pass(const ContView &cv, const double *wv, const double *vv, Epi &&epi) {
    if (cv.D == 4) cont_pass_d<4, MODE>(cv, wv, vv, epi);
    else cont_pass_d<6, MODE>(cv, wv, vv, epi);
}

__device__ __forceinline__ double cont_rowfac(const ContView &cv, int64_t n) {
    if (cv.D == 4) { double x[4]; cont_decode<4>(cv, n, x); return cont_const<4>(cv, x); }
    double x[6];
    cont_decode<6>(cv, n, x);
    return cont_const<6>(cv, x);
}
