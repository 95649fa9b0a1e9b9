"""Oracle SDF evaluation (test infrastructure only).  PARITY UNPINNED: the
reference contains no SDF code; this follows the Euler equation of
/root/reference/paper/autosdfs.tex:374-384 and the Epstein-Zin identity
-theta/psi + theta - 1 = -gamma:

  Mbar(n,n') = beta^theta * a_col(n') * e_sdf(n) * (w(n') / (w(n) - 1))^(theta-1)
  e_sdf(n)   = exp(-gamma (mu_c + z(n)) + 0.5 gamma^2 sigma_c(n)^2)
  q_f(n)     = sum_n' P(n,n') Mbar(n,n')                    (risk-free price)
  euler(n)   = beta^theta * s(n) / (w(n) - 1)^theta - 1,  s = a_row * P (a_col w^theta)

Independent pin (no shared derivation with the closed form above or with the CUDA
kernels): ``sdf_quadrature`` evaluates the UN-INTEGRATED one-period SDF of the paper,

  log M'(n,n',xi) = theta log beta + theta h_lam(n') - gamma (mu_c + z(n) + sigma_c(n) xi)
                    + (theta - 1) [log w(n') - log(w(n) - 1)],

state pair by state pair, and integrates the consumption shock xi ~ N(0,1) numerically
with a Gauss-Hermite rule: q_f(n) = sum_n' P(n,n') sum_q omega_q M'(n,n',xi_q) and the
pricing identity E[M' R_w'] = 1 with R_w' = (w(n')/(w(n)-1)) exp(mu_c + z(n) + sigma_c(n) xi).
The closed forms (e_sdf, the -gamma and +gamma^2/2 terms) are checked against it, and the
Euler identity evaluated this way involves e_sdf-type factors and the fixed point of the
reference-pinned operator T -- it is not the fixed-point residual in disguise.
"""
import numpy as np


def sdf_dense(w, P, a_row, a_col, e_sdf, β, θ):
    """Returns (q_f, euler_resid) for a dense operator."""
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    q_f = β ** θ * e_sdf * (w - 1) ** (1 - θ) * (P @ (a_col * w ** (θ - 1)))
    s = a_row * (P @ (a_col * w ** θ))
    euler = β ** θ * s / (w - 1) ** θ - 1
    return q_f, euler


def sdf_rows(w, P, a_col, e_sdf, β, θ, rows):
    """Explicit row tile of Mbar for the given current-state indices."""
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    rows = np.asarray(rows)
    ratio = w[None, :] / (w[rows, None] - 1)
    return β ** θ * a_col[None, :] * e_sdf[rows, None] * ratio ** (θ - 1)


def e_sdf_ssy(shapes, params, arrays):
    γ, μ_c = params[1], params[3]
    z, σ_c = arrays[6], arrays[8]
    e = np.exp(-γ * (μ_c + z))[None, None] * \
        np.exp(0.5 * (γ * σ_c) ** 2)[None, :, None, None]
    return np.broadcast_to(e, shapes).reshape(-1).copy()


def e_sdf_gcy(shapes, params, arrays):
    γ, μ_c = params[2], params[5]
    z, σ_c = arrays[0], arrays[9]
    zz = np.transpose(z, (3, 0, 1, 2))       # (z, zpi, hz, hzpi)
    e = np.exp(-γ * (μ_c + zz))[:, :, :, None, :, None] * \
        np.exp(0.5 * (γ * σ_c) ** 2)[None, None, None, :, None, None]
    return np.broadcast_to(e, shapes).reshape(-1).copy()


# --------------------------------------------------------------------------
# Independent check: un-integrated SDF + Gauss-Hermite quadrature over xi
# --------------------------------------------------------------------------
def state_fields_ssy(shapes, params, arrays):
    """Per-state (flattened C order) z(n), sigma_c(n), h_lam(n) of the SSY grid (l,k,i,j)."""
    h_λ, z, σ_c = arrays[0], arrays[6], arrays[8]
    zz = np.broadcast_to(z[None, None], shapes).reshape(-1)
    sc = np.broadcast_to(σ_c[None, :, None, None], shapes).reshape(-1)
    hl = np.broadcast_to(h_λ[:, None, None, None], shapes).reshape(-1)
    return zz, sc, hl


def state_fields_gcy(shapes, params, arrays):
    """Same for the GCY grid (z, z_pi, h_z, h_c, h_zpi, h_lam); z[i_zpi, i_hz, i_hzpi, i_z]."""
    z, σ_c, h_λ = arrays[0], arrays[9], arrays[13]
    zz = np.transpose(z, (3, 0, 1, 2))[:, :, :, None, :, None]
    zz = np.broadcast_to(zz, shapes).reshape(-1)
    sc = np.broadcast_to(σ_c[None, None, None, :, None, None], shapes).reshape(-1)
    hl = np.broadcast_to(h_λ[None, None, None, None, None, :], shapes).reshape(-1)
    return zz, sc, hl


def sdf_quadrature(w, P, z, σ_c, h_λ, β, γ, θ, μ_c, n_nodes=24):
    """(q_f, E[M' R_w'] - 1) from the un-integrated log M' by Gauss-Hermite quadrature.

    Everything is formed pair by pair as an explicit (N, N', Q) array -- O(N^2 Q), small
    grids only; none of e_sdf / a_row / a_col enters."""
    w = np.asarray(w, dtype=np.float64).reshape(-1)
    ξ, ω = np.polynomial.hermite_e.hermegauss(n_nodes)       # weight exp(-x^2/2)
    ω = ω / np.sqrt(2 * np.pi)
    g_c = μ_c + z[:, None, None] + σ_c[:, None, None] * ξ[None, None, :]          # (N, 1, Q)
    logM = (θ * np.log(β) + θ * h_λ[None, :, None] - γ * g_c
            + (θ - 1) * (np.log(w)[None, :, None] - np.log(w - 1)[:, None, None]))   # (N, N', Q)
    M = np.exp(logM)
    q_f = np.einsum('nm,nmq,q->n', P, M, ω)
    R_w = (w[None, :, None] / (w - 1)[:, None, None]) * np.exp(g_c)
    euler = np.einsum('nm,nmq,q->n', P, M * R_w, ω) - 1
    return q_f, euler
