// Device-side grid / transition builder.
//
// Replaces discretize_ssy (ssy/discrete/ssy_wc_ratio.py:23-79), discretize_gcy
// (gcy/discrete/gcy_wc_ratio.py:31-131) and the quantecon.rouwenhorst chains they
// call; expands the Markov factors into the dense single-index P of
// temp_ssy.py:106 and builds the diagonal scalings of temp_ssy.py:146.
//
// The Rouwenhorst kernels use only IEEE +,-,*,/,sqrt with explicit rounding
// intrinsics (no FMA contraction), so chain states and transition matrices are
// bit-identical to the NumPy evaluation of the same published recursion.
#include "common.cuh"

// One CTA per chain.  chain c: n states, persistence rho, innovation sd sigma[c],
// drift mu[c].  states -> states_out + c*n, matrix -> P_out + c*n*n.
// scratch: 2*n*n doubles per chain.
__global__ void k_rouwenhorst(int n, double rho, const double *__restrict__ sigma, double sigma_scalar,
                              const double *__restrict__ mu, double *__restrict__ states_out,
                              double *__restrict__ P_out, double *__restrict__ scratch) {
    const int c = blockIdx.x;
    const double sg = sigma ? sigma[c] : sigma_scalar;
    const double m = mu ? mu[c] : 0.0;
    double *st = states_out + (int64_t)c * n;
    double *A = scratch + (int64_t)c * 2 * n * n;
    double *B = A + (int64_t)n * n;
    const double p = __ddiv_rn(__dadd_rn(1.0, rho), 2.0);
    const double q = p;
    const double one_m_p = __dsub_rn(1.0, p), one_m_q = __dsub_rn(1.0, q);

    // states: linspace(-psi, psi, n) + mu/(1-rho), NumPy's evaluation order
    const double y_sd = __dsqrt_rn(__ddiv_rn(__dmul_rn(sg, sg), __dsub_rn(1.0, __dmul_rn(rho, rho))));
    const double psi = __dmul_rn(y_sd, __dsqrt_rn((double)(n - 1)));
    const double start = -psi, stop = psi;
    const double delta = __dsub_rn(stop, start);
    const double step = __ddiv_rn(delta, (double)(n - 1));
    const double shift = __ddiv_rn(m, __dsub_rn(1.0, rho));
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double y;
        if (step == 0.0) y = __dadd_rn(__dmul_rn(__ddiv_rn((double)i, (double)(n - 1)), delta), start);
        else y = __dadd_rn(__dmul_rn((double)i, step), start);
        if (i == n - 1) y = stop;
        st[i] = __dadd_rn(y, shift);
    }
    // matrix recursion
    if (threadIdx.x == 0) {
        A[0] = p; A[1] = one_m_p; A[2] = one_m_q; A[3] = q;   // 2x2, row stride 2
    }
    __syncthreads();
    double *cur = A, *nxt = B;
    for (int s = 3; s <= n; ++s) {
        const int o = s - 1;   // old size
        for (int e = threadIdx.x; e < s * s; e += blockDim.x) {
            const int i = e / s, j = e % s;
            double v = 0.0;
            if (i < o && j < o) v = __dadd_rn(v, __dmul_rn(p, cur[i * o + j]));
            if (i < o && j >= 1) v = __dadd_rn(v, __dmul_rn(one_m_p, cur[i * o + j - 1]));
            if (i >= 1 && j < o) v = __dadd_rn(v, __dmul_rn(one_m_q, cur[(i - 1) * o + j]));
            if (i >= 1 && j >= 1) v = __dadd_rn(v, __dmul_rn(q, cur[(i - 1) * o + j - 1]));
            if (i >= 1 && i <= s - 2) v = __ddiv_rn(v, 2.0);
            nxt[e] = v;
        }
        __syncthreads();
        double *t = cur; cur = nxt; nxt = t;
    }
    double *Po = P_out + (int64_t)c * n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Po[e] = cur[e];
}

__global__ void k_scale_exp(const double *__restrict__ h, double phi, double *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = phi * exp(h[i]);
}

// GCY z chain drifts: mu[(a*n_hz + b)*n_zpi + c] = rho_pi * zpi[a, c]; sigma likewise sigma_z[b]
__global__ void k_gcy_z_inputs(const double *__restrict__ zpi, const double *__restrict__ sigma_z,
                               double rho_pi, int n_hzpi, int n_hz, int n_zpi,
                               double *__restrict__ mu, double *__restrict__ sg) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = n_hzpi * n_hz * n_zpi;
    if (t >= total) return;
    const int c = t % n_zpi, b = (t / n_zpi) % n_hz, a = t / (n_zpi * n_hz);
    mu[t] = rho_pi * zpi[a * n_zpi + c];
    sg[t] = sigma_z[b];
}

// chains were produced in (a=h_zpi, b=h_z, c=z_pi) order; the reference stores
// z_states[c, b, a, :] and z_Q[c, b, a, :, :]
__global__ void k_gcy_z_permute(const double *__restrict__ st_in, const double *__restrict__ P_in,
                                int n_hzpi, int n_hz, int n_zpi, int n_z,
                                double *__restrict__ st_out, double *__restrict__ P_out) {
    const int chain = blockIdx.x;
    const int c = chain % n_zpi, b = (chain / n_zpi) % n_hz, a = chain / (n_zpi * n_hz);
    const int dst = (c * n_hz + b) * n_hzpi + a;
    for (int e = threadIdx.x; e < n_z; e += blockDim.x) st_out[(int64_t)dst * n_z + e] = st_in[(int64_t)chain * n_z + e];
    for (int e = threadIdx.x; e < n_z * n_z; e += blockDim.x)
        P_out[(int64_t)dst * n_z * n_z + e] = P_in[(int64_t)chain * n_z * n_z + e];
}

static int alloc_arr(sdfs_factors *f, int idx, int64_t n) {
    sdfs_ctx *ctx = f->ctx;
    CUDA_TRY(ctx, cudaMalloc(&f->d_arr[idx], (size_t)(n > 0 ? n : 1) * sizeof(double)));
    f->n_elems[idx] = n;
    return SDFS_OK;
}

static int run_chain(sdfs_ctx *ctx, int n, double rho, const double *d_sigma, double sigma_scalar,
                     const double *d_mu, int n_chains, double *d_states, double *d_P) {
    double *scratch = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&scratch, (size_t)n_chains * 2 * n * n * sizeof(double)));
    k_rouwenhorst<<<n_chains, 128, 0, ctx->stream>>>(n, rho, d_sigma, sigma_scalar, d_mu, d_states, d_P, scratch);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaFree(scratch));
    return SDFS_OK;
}

#define TRY(x) do { int _rc = (x); if (_rc != SDFS_OK) return _rc; } while (0)

static int64_t ssy_arr_size(const int *s, int idx) {
    const int64_t L = s[0], K = s[1], I = s[2], J = s[3];
    const int64_t sz[10] = {L, L * L, K, K * K, I, I * I, I * J, I * J * J, K, I};
    return sz[idx];
}
static int64_t gcy_arr_size(const int *s, int idx) {
    const int64_t nz = s[0], nzp = s[1], nhz = s[2], nhc = s[3], nhzp = s[4], nhl = s[5];
    const int64_t sz[15] = {nzp * nhz * nhzp * nz, nzp * nhz * nhzp * nz * nz, nhzp * nzp, nhzp * nzp * nzp,
                            nhz, nhz * nhz, nhz, nhc, nhc * nhc, nhc, nhzp, nhzp * nhzp, nhzp, nhl, nhl * nhl};
    return sz[idx];
}

extern "C" {

int sdfs_factors_destroy(sdfs_factors *f) {
    if (!f) return SDFS_OK;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    for (int i = 0; i < 16; ++i)
        if (f->d_arr[i]) cudaFree(f->d_arr[i]);
    delete f;
    return SDFS_OK;
}

int sdfs_factors_count(sdfs_factors *f, int *n_arrays) {
    if (!f || !n_arrays) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_factors_count: NULL");
    *n_arrays = f->n_arrays;
    return SDFS_OK;
}

int sdfs_factors_array(sdfs_factors *f, int idx, int64_t *n_elems, const double **d_ptr) {
    if (!f) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_factors_array: NULL");
    ARG_CHECK(f->ctx, idx >= 0 && idx < f->n_arrays);
    if (n_elems) *n_elems = f->n_elems[idx];
    if (d_ptr) *d_ptr = f->d_arr[idx];
    return SDFS_OK;
}

static int factors_new(sdfs_ctx *ctx, int model, const double *h_params, const int32_t *h_shapes,
                       sdfs_factors **out) {
    ARG_CHECK(ctx, ctx && h_params && h_shapes && out);
    ARG_CHECK(ctx, model == SDFS_MODEL_SSY || model == SDFS_MODEL_GCY);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    sdfs_factors *f = new sdfs_factors();
    f->ctx = ctx;
    f->model = model;
    f->D = (model == SDFS_MODEL_SSY) ? 4 : 6;
    const int np = (model == SDFS_MODEL_SSY) ? 13 : 18;
    for (int i = 0; i < np; ++i) f->params[i] = h_params[i];
    for (int d = 0; d < f->D; ++d) {
        if (h_shapes[d] < 2 || h_shapes[d] > 4096) {
            delete f;
            return sdfs_set_error(ctx, SDFS_ERR_ARG, "shapes[%d]=%d: every axis needs 2..4096 states", d, h_shapes[d]);
        }
        f->shapes[d] = h_shapes[d];
    }
    f->n_arrays = (model == SDFS_MODEL_SSY) ? 10 : 15;
    for (int i = 0; i < f->n_arrays; ++i) {
        const int64_t n = (model == SDFS_MODEL_SSY) ? ssy_arr_size(f->shapes, i) : gcy_arr_size(f->shapes, i);
        int rc = alloc_arr(f, i, n);
        if (rc != SDFS_OK) { sdfs_factors_destroy(f); return rc; }
    }
    *out = f;
    return SDFS_OK;
}

int sdfs_factors_from_host(sdfs_ctx *ctx, int model, const double *h_params, const int32_t *h_shapes,
                           const double *const *h_arrays, int n_arrays, sdfs_factors **out) {
    sdfs_factors *f = nullptr;
    TRY(factors_new(ctx, model, h_params, h_shapes, &f));
    if (n_arrays != f->n_arrays || !h_arrays) {
        sdfs_factors_destroy(f);
        return sdfs_set_error(ctx, SDFS_ERR_ARG, "expected %d factor arrays, got %d", f->n_arrays, n_arrays);
    }
    for (int i = 0; i < n_arrays; ++i) {
        cudaError_t e = cudaMemcpyAsync(f->d_arr[i], h_arrays[i], (size_t)f->n_elems[i] * sizeof(double),
                                        cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { sdfs_factors_destroy(f); return sdfs_set_error(ctx, SDFS_ERR_CUDA, "factor upload: %s", cudaGetErrorString(e)); }
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *out = f;
    return SDFS_OK;
}

int sdfs_factors_build(sdfs_ctx *ctx, int model, const double *h_params, const int32_t *h_shapes,
                       sdfs_factors **out) {
    sdfs_factors *f = nullptr;
    TRY(factors_new(ctx, model, h_params, h_shapes, &f));
    const double *pr = f->params;
    int rc = SDFS_OK;
    auto scale_exp = [&](const double *h, double phi, double *o, int n) {
        k_scale_exp<<<(n + 127) / 128, 128, 0, ctx->stream>>>(h, phi, o, n);
        ctx->launches++;
    };
    if (model == SDFS_MODEL_SSY) {
        // params: beta,gamma,psi,mu_c,rho,phi_z,phi_c,rho_z,rho_c,rho_lam,s_z,s_c,s_lam
        const int L = f->shapes[0], K = f->shapes[1], I = f->shapes[2], J = f->shapes[3];
        const double rho = pr[4], phi_z = pr[5], phi_c = pr[6], rho_z = pr[7], rho_c = pr[8], rho_l = pr[9];
        const double s_z = pr[10], s_c = pr[11], s_l = pr[12];
        // arrays: h_lam,Q_lam,h_c,Q_c,h_z,Q_hz,z,z_Q,sigma_c,sigma_z
        if ((rc = run_chain(ctx, L, rho_l, nullptr, s_l, nullptr, 1, f->d_arr[0], f->d_arr[1]))) goto fail;
        if ((rc = run_chain(ctx, K, rho_c, nullptr, s_c, nullptr, 1, f->d_arr[2], f->d_arr[3]))) goto fail;
        if ((rc = run_chain(ctx, I, rho_z, nullptr, s_z, nullptr, 1, f->d_arr[4], f->d_arr[5]))) goto fail;
        scale_exp(f->d_arr[2], phi_c, f->d_arr[8], K);
        scale_exp(f->d_arr[4], phi_z, f->d_arr[9], I);
        if ((rc = run_chain(ctx, J, rho, f->d_arr[9], 0.0, nullptr, I, f->d_arr[6], f->d_arr[7]))) goto fail;
    } else {
        // params: beta,psi,gamma,rho_lam,s_lam,mu_c,phi_c,rho,rho_pi,phi_z,rho_c,s_c,rho_z,s_z,
        //         rho_pipi,phi_zpi,rho_zpi,s_zpi
        const int nz = f->shapes[0], nzp = f->shapes[1], nhz = f->shapes[2], nhc = f->shapes[3],
                  nhzp = f->shapes[4], nhl = f->shapes[5];
        const double rho_l = pr[3], s_l = pr[4], phi_c = pr[6], rho = pr[7], rho_pi = pr[8], phi_z = pr[9],
                     rho_c = pr[10], s_c = pr[11], rho_z = pr[12], s_z = pr[13], rho_pp = pr[14],
                     phi_zp = pr[15], rho_zp = pr[16], s_zp = pr[17];
        // arrays: z,z_Q,zpi,zpi_Q,h_z,Q_hz,sig_z,h_c,Q_hc,sig_c,h_zpi,Q_hzpi,sig_zpi,h_lam,Q_hlam
        if ((rc = run_chain(ctx, nhz, rho_z, nullptr, s_z, nullptr, 1, f->d_arr[4], f->d_arr[5]))) goto fail;
        if ((rc = run_chain(ctx, nhc, rho_c, nullptr, s_c, nullptr, 1, f->d_arr[7], f->d_arr[8]))) goto fail;
        if ((rc = run_chain(ctx, nhzp, rho_zp, nullptr, s_zp, nullptr, 1, f->d_arr[10], f->d_arr[11]))) goto fail;
        if ((rc = run_chain(ctx, nhl, rho_l, nullptr, s_l, nullptr, 1, f->d_arr[13], f->d_arr[14]))) goto fail;
        scale_exp(f->d_arr[4], phi_z, f->d_arr[6], nhz);
        scale_exp(f->d_arr[7], phi_c, f->d_arr[9], nhc);
        scale_exp(f->d_arr[10], phi_zp, f->d_arr[12], nhzp);
        if ((rc = run_chain(ctx, nzp, rho_pp, f->d_arr[12], 0.0, nullptr, nhzp, f->d_arr[2], f->d_arr[3]))) goto fail;
        const int chains = nhzp * nhz * nzp;
        double *mu = nullptr, *sg = nullptr, *st = nullptr, *PQ = nullptr;
        cudaMalloc(&mu, chains * sizeof(double));
        cudaMalloc(&sg, chains * sizeof(double));
        cudaMalloc(&st, (size_t)chains * nz * sizeof(double));
        cudaError_t e = cudaMalloc(&PQ, (size_t)chains * nz * nz * sizeof(double));
        if (e != cudaSuccess) { rc = sdfs_set_error(ctx, SDFS_ERR_NOMEM, "gcy builder scratch: %s", cudaGetErrorString(e)); goto fail; }
        k_gcy_z_inputs<<<(chains + 127) / 128, 128, 0, ctx->stream>>>(f->d_arr[2], f->d_arr[6], rho_pi, nhzp, nhz, nzp, mu, sg);
        ctx->launches++;
        rc = run_chain(ctx, nz, rho, sg, 0.0, mu, chains, st, PQ);
        if (rc == SDFS_OK) {
            k_gcy_z_permute<<<chains, 128, 0, ctx->stream>>>(st, PQ, nhzp, nhz, nzp, nz, f->d_arr[0], f->d_arr[1]);
            ctx->launches++;
            cudaStreamSynchronize(ctx->stream);
        }
        cudaFree(mu); cudaFree(sg); cudaFree(st); cudaFree(PQ);
        if (rc) goto fail;
    }
    {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = sdfs_set_error(ctx, SDFS_ERR_CUDA, "factor build: %s", cudaGetErrorString(e)); goto fail; }
    }
    *out = f;
    return SDFS_OK;
fail:
    sdfs_factors_destroy(f);
    return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Mode description shared by the factor-form apply and the dense expansion.
// ---------------------------------------------------------------------------
#define KRON_TC_MIN_HOST 9      // = KRON_TC_MIN of rowdot.cuh (shortest axis on the tensor-core contraction)
int factors_to_kron(const sdfs_factors *f, KronView *kv) {
    memset(kv, 0, sizeof(*kv));
    kv->D = f->D;
    kv->N = 1;
    for (int d = 0; d < f->D; ++d) { kv->shape[d] = f->shapes[d]; kv->N *= f->shapes[d]; }
    if (f->model == SDFS_MODEL_SSY) {
        // contraction order of the sum-factorised apply: l' (Q_lam) first -- the leading axis is the one a
        // multi-GPU run splits into slabs, and contracting it first is the only step that needs the other
        // ranks' part of the input, so everything after it is rank-local (the same order on one GPU keeps
        // results bit-identical across rank counts) -- then k' (Q_c), i' (Q_hz) and last j' (z_Q[i], which needs
        // the current i): the last contraction runs along the innermost axis, so a finished fibre is one
        // contiguous run of outputs (wide stores in the fused epilogue, also into the peers' result buffers)
        kv->n_modes = 4;
        kv->modes[0].mat = f->d_arr[1]; kv->modes[0].dim = 0;
        kv->modes[1].mat = f->d_arr[3]; kv->modes[1].dim = 1;
        kv->modes[2].mat = f->d_arr[5]; kv->modes[2].dim = 2;
        kv->modes[3].mat = f->d_arr[7]; kv->modes[3].dim = 3; kv->modes[3].mstride[2] = 1;
    } else {
        const int nhz = f->shapes[2], nhzp = f->shapes[4];
        kv->n_modes = 6;
        kv->modes[0].mat = f->d_arr[14]; kv->modes[0].dim = 5;                 // h_lam
        kv->modes[1].mat = f->d_arr[11]; kv->modes[1].dim = 4;                 // h_zpi
        kv->modes[2].mat = f->d_arr[8];  kv->modes[2].dim = 3;                 // h_c
        kv->modes[3].mat = f->d_arr[5];  kv->modes[3].dim = 2;                 // h_z
        kv->modes[4].mat = f->d_arr[3];  kv->modes[4].dim = 1;                 // z_pi: zpi_Q[i_hzpi]
        kv->modes[4].mstride[4] = 1;
        kv->modes[5].mat = f->d_arr[1];  kv->modes[5].dim = 0;                 // z: z_Q[i_zpi,i_hz,i_hzpi]
        kv->modes[5].mstride[1] = nhz * nhzp; kv->modes[5].mstride[2] = nhzp; kv->modes[5].mstride[4] = 1;
    }
    // fibre-kernel work decomposition per mode
    long long estride[SDFS_MAX_DIMS];
    estride[kv->D - 1] = 1;
    for (int d = kv->D - 2; d >= 0; --d) estride[d] = estride[d + 1] * kv->shape[d + 1];
    for (int m = 0; m < kv->n_modes; ++m) {
        KronMode &md = kv->modes[m];
        md.stride = estride[md.dim];
        md.nF = md.nM = 0;
        md.Fcount = md.Mcount = 1;
        for (int d = 0; d < kv->D; ++d) {
            if (d == md.dim) continue;
            if (md.mstride[d] != 0) {
                md.Mshape[md.nM] = kv->shape[d]; md.Mstride[md.nM] = estride[d]; md.Mmat[md.nM] = md.mstride[d];
                md.Mcount *= kv->shape[d]; md.nM++;
            } else {
                md.Fshape[md.nF] = kv->shape[d]; md.Fstride[md.nF] = estride[d];
                md.Fcount *= kv->shape[d]; md.nF++;
            }
        }
        md.out0 = 0; md.nout = kv->shape[md.dim]; md.base_off = 0; md.colscale = nullptr;
    }
    kv->lead0 = 0; kv->leadn = kv->shape[0];
    kv->row_begin = 0; kv->row_end = kv->N;
    return SDFS_OK;
}

// Can the leading axis of this view be split into per-rank slabs?  The mode that contracts axis 0 must come
// first (its matrix cannot depend on another coordinate, and no other matrix may depend on coordinate 0) and
// run on the tensor-core path, which has the restricted-output variant.  True for SSY (axis 0 = h_lambda);
// false for GCY, whose leading axis z is contracted last by matrices indexed by (z_pi, h_z, h_zpi).
bool kron_can_shard(const KronView &kv) {
    const KronMode &m0 = kv.modes[0];
    if (m0.dim != 0 || m0.nM != 0) return false;
    for (int m = 1; m < kv.n_modes; ++m)
        if (kv.modes[m].mstride[0] != 0) return false;
    return kv.shape[0] >= KRON_TC_MIN_HOST && kv.shape[0] <= 64;
}

// Restrict a view to the slab [l0, l1) of axis 0: rows [l0, l1) x (everything else) of the result.
void kron_restrict_leading(KronView *kv, int l0, int l1) {
    const long long inner = kv->N / kv->shape[0];
    kv->lead0 = l0; kv->leadn = l1 - l0;
    kv->row_begin = (int64_t)l0 * inner; kv->row_end = (int64_t)l1 * inner;
    KronMode &m0 = kv->modes[0];
    m0.out0 = l0; m0.nout = l1 - l0;
    for (int m = 1; m < kv->n_modes; ++m) {
        KronMode &md = kv->modes[m];
        md.base_off = (long long)l0 * inner;
        for (int a = 0; a < md.nF; ++a)
            if (md.Fstride[a] == inner) {          // axis 0 comes first in the free-axis list and is the slowest axis
                md.Fcount = md.Fcount / md.Fshape[a] * (l1 - l0);
                md.Fshape[a] = l1 - l0;
                break;
            }
    }
}

// a_row, a_col, e_sdf for every state (C-order flattening, temp_ssy.py:41-42).
__global__ void k_build_scalings(int model, KronView kv, const double *__restrict__ h_lam,
                                 const double *__restrict__ sigma_c, const double *__restrict__ z,
                                 double gamma, double theta, double mu_c, double *__restrict__ a_row,
                                 double *__restrict__ a_col, double *__restrict__ e_sdf) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < kv.N; n += (int64_t)gridDim.x * blockDim.x) {
        int c[SDFS_MAX_DIMS];
        int64_t rem = n;
        for (int d = kv.D - 1; d >= 0; --d) { c[d] = (int)(rem % kv.shape[d]); rem /= kv.shape[d]; }
        double hl, sc, zz;
        if (model == SDFS_MODEL_SSY) {         // (l,k,i,j); z[i,j]
            hl = h_lam[c[0]]; sc = sigma_c[c[1]]; zz = z[c[2] * kv.shape[3] + c[3]];
        } else {                                // (z,zpi,hz,hc,hzpi,hlam); z[i_zpi,i_hz,i_hzpi,i_z]
            hl = h_lam[c[5]]; sc = sigma_c[c[3]];
            zz = z[((c[1] * kv.shape[2] + c[2]) * kv.shape[4] + c[4]) * kv.shape[0] + c[0]];
        }
        const double omg = 1.0 - gamma;
        const double t2 = omg * sc;
        a_col[n] = exp(theta * hl);
        a_row[n] = exp(0.5 * (t2 * t2)) * exp(omg * (mu_c + zz));
        const double g2 = gamma * sc;
        e_sdf[n] = exp(-gamma * (mu_c + zz)) * exp(0.5 * (g2 * g2));
    }
}

// Dense expansion: P[n, n'] = prod_m M_m[mat_m(n)][i_m(n)][i_m(n')] for local rows.
// One CTA per row.  The column multi-index is split into a leading part (axes 0..h-1, A
// combinations) and a trailing part (axes h..D-1, B combinations, A*B = N); the partial
// products over each part are tabulated in shared memory once per row, so every element of
// the row costs one multiply and the kernel runs at the HBM write rate instead of being
// bound by 2 D integer divisions per element.
__global__ void __launch_bounds__(256) k_expand_dense(KronView kv, int64_t row_begin, int64_t row_end, int64_t ld,
                                                      int h, int A, int B, double *__restrict__ P) {
    extern __shared__ double sh[];
    double *srow = sh;                         // concatenated factor rows of this P row
    __shared__ int soff[SDFS_MAX_DIMS];        // offset of the factor row of axis d in srow
    int nsum = 0;
    for (int d = 0; d < kv.D; ++d) nsum += kv.shape[d];
    double *Ta = sh + nsum, *Tb = Ta + A;
    for (int64_t row = row_begin + blockIdx.x; row < row_end; row += gridDim.x) {
        __syncthreads();
        int c[SDFS_MAX_DIMS];
        int64_t rem = row;
        for (int d = kv.D - 1; d >= 0; --d) { c[d] = (int)(rem % kv.shape[d]); rem /= kv.shape[d]; }
        int off = 0;
        for (int d = 0; d < kv.D; ++d) {
            // the mode that contracts axis d
            int m = 0;
            for (int mm = 0; mm < kv.n_modes; ++mm) if (kv.modes[mm].dim == d) m = mm;
            const KronMode &md = kv.modes[m];
            const int n = kv.shape[d];
            int mat = 0;
            for (int dd = 0; dd < kv.D; ++dd) mat += c[dd] * md.mstride[dd];
            const double *src = md.mat + ((int64_t)mat * n + c[d]) * n;
            for (int j = threadIdx.x; j < n; j += blockDim.x) srow[off + j] = src[j];
            if (threadIdx.x == 0) soff[d] = off;
            off += n;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < A + B; e += blockDim.x) {
            const bool lead = e < A;
            int r2 = lead ? e : e - A;
            double v = 1.0;
            // multiply in increasing-axis order within each part
            const int d0 = lead ? 0 : h, d1 = lead ? h : kv.D;
            int cc[SDFS_MAX_DIMS];
            for (int d = d1 - 1; d >= d0; --d) { cc[d] = r2 % kv.shape[d]; r2 /= kv.shape[d]; }
            for (int d = d0; d < d1; ++d) v *= srow[soff[d] + cc[d]];
            if (lead) Ta[e] = v; else Tb[e - A] = v;
        }
        __syncthreads();
        double *dst = P + (row - row_begin) * ld;
        // two columns per thread and step (16-byte stores); ld is even and >= N
        for (int64_t col = 2 * (int64_t)threadIdx.x; col < ld; col += 2 * blockDim.x) {
            double2 v = make_double2(0.0, 0.0);
            if (col < kv.N) {
                const int ia = (int)(col / B), ib = (int)(col - (int64_t)ia * B);
                v.x = Ta[ia] * Tb[ib];
                if (col + 1 < kv.N) v.y = (ib + 1 < B) ? Ta[ia] * Tb[ib + 1] : Ta[ia + 1] * Tb[0];
            }
            *reinterpret_cast<double2 *>(dst + col) = v;
        }
    }
}

// wc_loglinear on the grid (ssy_model.py:143-153, gcy_model.py:146-157)
__global__ void k_loglinear(int model, KronView kv, const double *__restrict__ h_lam, const double *__restrict__ h_c,
                            const double *__restrict__ h_z, const double *__restrict__ z, const double *__restrict__ h_zpi,
                            const double *__restrict__ zpi, double phi_c, double phi_z, double phi_zpi, double A0,
                            double Al, double Ac, double Ahz, double Az, double Ahzpi, double Azpi, int exponentiate,
                            double *__restrict__ out) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < kv.N; n += (int64_t)gridDim.x * blockDim.x) {
        int c[SDFS_MAX_DIMS];
        int64_t rem = n;
        for (int d = kv.D - 1; d >= 0; --d) { c[d] = (int)(rem % kv.shape[d]); rem /= kv.shape[d]; }
        double v;
        if (model == SDFS_MODEL_SSY) {          // (l,k,i,j)
            const double sz = h_z[c[2]] * 2.0 * (phi_z * phi_z) + phi_z * phi_z;
            const double sc = h_c[c[1]] * 2.0 * (phi_c * phi_c) + phi_c * phi_c;
            v = A0 + Al * h_lam[c[0]] + Ac * sc + Ahz * sz + Az * z[c[2] * kv.shape[3] + c[3]];
        } else {                                 // (z,zpi,hz,hc,hzpi,hlam)
            const double sz = h_z[c[2]] * 2.0 * (phi_z * phi_z) + phi_z * phi_z;
            const double sc = h_c[c[3]] * 2.0 * (phi_c * phi_c) + phi_c * phi_c;
            const double sp = h_zpi[c[4]] * 2.0 * (phi_zpi * phi_zpi) + phi_zpi * phi_zpi;
            const double zz = z[((c[1] * kv.shape[2] + c[2]) * kv.shape[4] + c[4]) * kv.shape[0] + c[0]];
            const double zp = zpi[c[4] * kv.shape[1] + c[1]];
            v = A0 + Al * h_lam[c[5]] + Ac * sc + Ahz * sz + Az * zz + Ahzpi * sp + Azpi * zp;
        }
        out[n] = exponentiate ? exp(v) : v;
    }
}

extern "C" int sdfs_factors_loglinear(sdfs_factors *f, const double *h_coeffs, int exponentiate, double *d_out) {
    if (!f) return sdfs_set_error(nullptr, SDFS_ERR_ARG, "sdfs_factors_loglinear: NULL");
    sdfs_ctx *ctx = f->ctx;
    ARG_CHECK(ctx, h_coeffs && d_out);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    KronView kv;
    factors_to_kron(f, &kv);
    const double *pr = f->params;
    const int grid = (int)((kv.N + 255) / 256 < 4096 ? (kv.N + 255) / 256 : 4096);
    if (f->model == SDFS_MODEL_SSY)
        k_loglinear<<<grid, 256, 0, ctx->stream>>>(f->model, kv, f->d_arr[0], f->d_arr[2], f->d_arr[4], f->d_arr[6], nullptr, nullptr,
                                                   pr[6], pr[5], 0.0, h_coeffs[0], h_coeffs[1], h_coeffs[2], h_coeffs[3], h_coeffs[4],
                                                   0.0, 0.0, exponentiate, d_out);
    else
        k_loglinear<<<grid, 256, 0, ctx->stream>>>(f->model, kv, f->d_arr[13], f->d_arr[7], f->d_arr[4], f->d_arr[0], f->d_arr[10],
                                                   f->d_arr[2], pr[6], pr[9], pr[15], h_coeffs[0], h_coeffs[1], h_coeffs[2],
                                                   h_coeffs[3], h_coeffs[4], h_coeffs[5], h_coeffs[6], exponentiate, d_out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

int launch_build_scalings(sdfs_ctx *ctx, const sdfs_factors *f, const KronView &kv, double gamma,
                          double theta, double mu_c, double *a_row, double *a_col, double *e_sdf) {
    const double *h_lam = (f->model == SDFS_MODEL_SSY) ? f->d_arr[0] : f->d_arr[13];
    const double *sig_c = (f->model == SDFS_MODEL_SSY) ? f->d_arr[8] : f->d_arr[9];
    const double *z = (f->model == SDFS_MODEL_SSY) ? f->d_arr[6] : f->d_arr[0];
    const int grid = (int)((kv.N + 255) / 256 < 4096 ? (kv.N + 255) / 256 : 4096);
    k_build_scalings<<<grid, 256, 0, ctx->stream>>>(f->model, kv, h_lam, sig_c, z, gamma, theta, mu_c, a_row, a_col, e_sdf);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}

int launch_expand_dense(sdfs_ctx *ctx, const KronView &kv, int64_t row_begin, int64_t row_end, int64_t ld, double *P) {
    int nsum = 0;
    for (int d = 0; d < kv.D; ++d) nsum += kv.shape[d];
    // split point: leading part as close to sqrt(N) as possible
    int h = 1, best_h = 1;
    double best = 1e300;
    for (h = 1; h < kv.D; ++h) {
        double a = 1, b = 1;
        for (int d = 0; d < h; ++d) a *= kv.shape[d];
        for (int d = h; d < kv.D; ++d) b *= kv.shape[d];
        if (a + b < best) { best = a + b; best_h = h; }
    }
    h = best_h;
    int64_t A = 1, B = 1;
    for (int d = 0; d < h; ++d) A *= kv.shape[d];
    for (int d = h; d < kv.D; ++d) B *= kv.shape[d];
    const size_t smem = (size_t)(nsum + A + B) * sizeof(double);
    if (smem > 200 * 1024)
        return sdfs_set_error(ctx, SDFS_ERR_UNSUPPORTED, "dense expansion tables need %zu bytes of shared memory", smem);
    const int64_t rows = row_end - row_begin;
    const int grid = (int)(rows < (int64_t)ctx->sm_count * 8 ? rows : (int64_t)ctx->sm_count * 8);
    if (grid <= 0) return SDFS_OK;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_expand_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_expand_dense<<<grid, 256, smem, ctx->stream>>>(kv, row_begin, row_end, ld, h, (int)A, (int)B, P);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SDFS_OK;
}
