"""Oracle restatement of the log-linear closed form (test infrastructure only).

Follows /root/reference/code/ssy/ssy_model.py:86-156 and
/root/reference/code/gcy/gcy_model.py:80-159 (wc_loglinear_factory): a scalar root
(scipy brentq on [-20, 20], as in the reference) fixes the constants, then log w is affine
in the state.  PINNED against the importable reference files through
tests/golden/loglinear.json (tests/golden/make_golden.py).
"""
import numpy as np
from scipy.optimize import brentq


def _common(β, ψ, θ, μ_c, ρ, ρ_λ, s_λ, φ_c, s_c, ρ_c, φ_z, s_z, ρ_z, extra=None):
    s_wc = 2 * φ_c ** 2 * s_c
    s_wx = 2 * φ_z ** 2 * s_z
    k1 = lambda x: np.exp(x) / (1 + np.exp(x))
    k0 = lambda x: np.log(1 + np.exp(x)) - k1(x) * x
    A1 = lambda x: (1 - 1 / ψ) / (1 - k1(x) * ρ)
    Aλ = lambda x: ρ_λ / (1 - k1(x) * ρ_λ)
    Az = lambda x: (θ / 2) * (k1(x) * A1(x)) ** 2 / (1 - k1(x) * ρ_z)
    Ac = lambda x: (θ / 2) * (1 - 1 / ψ) ** 2 / (1 - k1(x) * ρ_c)
    if extra is None:
        Aπ = Azπ = None
        lin = lambda x: 0.0
        quad = lambda x: 0.0
        sub = lambda x: 0.0
    else:
        ρ_π, ρ_ππ, φ_zπ, s_zπ, ρ_zπ = extra
        s_wxπ = 2 * φ_zπ ** 2 * s_zπ
        Aπ = lambda x: k1(x) * (1 - 1 / ψ) * ρ_π / ((1 - k1(x) * ρ) * (1 - k1(x) * ρ_ππ))
        Azπ = lambda x: (θ / 2) * (k1(x) * Aπ(x)) ** 2 / (1 - k1(x) * ρ_zπ)
        lin = lambda x: k1(x) * Azπ(x) * φ_zπ ** 2 * (1 - ρ_zπ)
        quad = lambda x: (k1(x) * Azπ(x) * s_wxπ) ** 2
        sub = lambda x: Azπ(x) * φ_zπ ** 2
    A0 = lambda x: (np.log(β) + k0(x) + μ_c * (1 - 1 / ψ)
                    + k1(x) * Az(x) * φ_z ** 2 * (1 - ρ_z)
                    + k1(x) * Ac(x) * φ_c ** 2 * (1 - ρ_c) + lin(x)
                    + (θ / 2) * ((k1(x) * Aλ(x) + 1) ** 2 * s_λ ** 2 + (k1(x) * Az(x) * s_wx) ** 2
                                 + (k1(x) * Ac(x) * s_wc) ** 2 + quad(x))) / (1 - k1(x))
    q = brentq(lambda x: x - A0(x) - Ac(x) * φ_c ** 2 - Az(x) * φ_z ** 2 - sub(x), -20, 20)
    out = dict(A0=A0(q), Ah_λ=Aλ(q), Ah_c=Ac(q), Ah_z=Az(q), Az=A1(q), qbar=q)
    if extra is not None:
        out.update(Ah_zπ=Azπ(q), Az_π=Aπ(q))
    return out


def loglinear_ssy(ssy):
    β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ = ssy.params
    θ = (1 - γ) / (1 - 1 / ψ)
    c = _common(β, ψ, θ, μ_c, ρ, ρ_λ, s_λ, ϕ_c, s_c, ρ_c, ϕ_z, s_z, ρ_z)

    def f(x):
        h_λ, h_c, h_z, z = x
        return (c["A0"] + c["Ah_λ"] * h_λ + c["Ah_c"] * (h_c * 2 * ϕ_c ** 2 + ϕ_c ** 2)
                + c["Ah_z"] * (h_z * 2 * ϕ_z ** 2 + ϕ_z ** 2) + c["Az"] * z)
    f.coeffs = c
    return f


def loglinear_gcy(gcy):
    (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z, ρ_ππ, φ_zπ, ρ_zπ, s_zπ) = gcy.params
    θ = (1 - γ) / (1 - 1 / ψ)
    c = _common(β, ψ, θ, μ_c, ρ, ρ_λ, s_λ, φ_c, s_c, ρ_c, φ_z, s_z, ρ_z, extra=(ρ_π, ρ_ππ, φ_zπ, s_zπ, ρ_zπ))

    def f(x):
        h_λ, h_c, h_z, h_zπ, z, z_π = x
        return (c["A0"] + c["Ah_λ"] * h_λ + c["Ah_c"] * (h_c * 2 * φ_c ** 2 + φ_c ** 2)
                + c["Ah_z"] * (h_z * 2 * φ_z ** 2 + φ_z ** 2) + c["Az"] * z
                + c["Ah_zπ"] * (h_zπ * 2 * φ_zπ ** 2 + φ_zπ ** 2) + c["Az_π"] * z_π)
    f.coeffs = c
    return f


def loglinear_grid_ssy(ssy, shapes, arrays):
    """log w of the closed form on the discretised grid, state order (l, k, i, j)."""
    f = loglinear_ssy(ssy)
    h_λ, _, h_c, _, h_z, _, z = arrays[:7]
    L, K, I, J = shapes
    out = np.empty(shapes)
    for l in range(L):
        for k in range(K):
            for i in range(I):
                for j in range(J):
                    out[l, k, i, j] = f((h_λ[l], h_c[k], h_z[i], z[i, j]))
    return out


def loglinear_grid_gcy(gcy, shapes, arrays):
    """state order (z, z_pi, h_z, h_c, h_zpi, h_lam); z[i_zpi,i_hz,i_hzpi,i_z], zpi[i_hzpi,i_zpi]."""
    f = loglinear_gcy(gcy)
    z, _, zπ, _, h_z, _, _, h_c, _, _, h_zπ, _, _, h_λ, _ = arrays
    out = np.empty(shapes)
    for idx in np.ndindex(*shapes):
        iz, izp, ihz, ihc, ihzp, ihl = idx
        out[idx] = f((h_λ[ihl], h_c[ihc], h_z[ihz], h_zπ[ihzp], z[izp, ihz, ihzp, iz], zπ[ihzp, izp]))
    return out
