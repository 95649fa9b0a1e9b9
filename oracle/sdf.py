"""Oracle SDF evaluation (test infrastructure only).  PARITY UNPINNED: the
reference contains no SDF code; this follows the Euler equation of
/root/reference/paper/autosdfs.tex:374-384 and the Epstein-Zin identity
-theta/psi + theta - 1 = -gamma:

  Mbar(n,n') = beta^theta * a_col(n') * e_sdf(n) * (w(n') / (w(n) - 1))^(theta-1)
  e_sdf(n)   = exp(-gamma (mu_c + z(n)) + 0.5 gamma^2 sigma_c(n)^2)
  q_f(n)     = sum_n' P(n,n') Mbar(n,n')                    (risk-free price)
  euler(n)   = beta^theta * s(n) / (w(n) - 1)^theta - 1,  s = a_row * P (a_col w^theta)

The only available check is euler ~ 0 at the fixed point (E[M R_w] = 1).
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
