"""Oracle discretisers (test infrastructure only).

discretize_ssy follows /root/reference/code/ssy/discrete/ssy_wc_ratio.py:23-79,
discretize_gcy follows /root/reference/code/gcy/discrete/gcy_wc_ratio.py:31-131.
Return tuples keep the reference order so the reference's own loop oracles can
be fed with them (tests/golden/make_golden.py).
"""
import numpy as np
from .rouwenhorst import rouwenhorst


def discretize_ssy(ssy, shapes):
    L, K, I, J = shapes
    (β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ) = ssy.params
    h_λ, Q_λ = rouwenhorst(L, ρ_λ, s_λ, 0)
    h_c, Q_c = rouwenhorst(K, ρ_c, s_c, 0)
    h_z, Q_hz = rouwenhorst(I, ρ_z, s_z, 0)
    σ_z = ϕ_z * np.exp(h_z)
    σ_c = ϕ_c * np.exp(h_c)
    z = np.empty((I, J))
    z_Q = np.empty((I, J, J))
    for i in range(I):
        z[i], z_Q[i] = rouwenhorst(J, ρ, σ_z[i], 0)
    return (h_λ, Q_λ, h_c, Q_c, h_z, Q_hz, z, z_Q, σ_c, σ_z)


def discretize_gcy(gcy, shapes):
    n_z, n_zπ, n_hz, n_hc, n_hzπ, n_hλ = shapes
    (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z,
     ρ_ππ, φ_zπ, ρ_zπ, s_zπ) = gcy.params
    h_z, Q_hz = rouwenhorst(n_hz, ρ_z, s_z)
    h_c, Q_hc = rouwenhorst(n_hc, ρ_c, s_c)
    h_zπ, Q_hzπ = rouwenhorst(n_hzπ, ρ_zπ, s_zπ)
    h_λ, Q_hλ = rouwenhorst(n_hλ, ρ_λ, s_λ)
    σ_z = φ_z * np.exp(h_z)
    σ_c = φ_c * np.exp(h_c)
    σ_zπ = φ_zπ * np.exp(h_zπ)
    # z_pi chain: one grid per h_zpi state
    zπ = np.empty((n_hzπ, n_zπ))
    zπ_Q = np.empty((n_hzπ, n_zπ, n_zπ))
    for a in range(n_hzπ):
        zπ[a], zπ_Q[a] = rouwenhorst(n_zπ, ρ_ππ, σ_zπ[a])
    # z chain: one grid per (z_pi, h_z, h_zpi), drift rho_pi * z_pi
    z = np.empty((n_zπ, n_hz, n_hzπ, n_z))
    z_Q = np.empty((n_zπ, n_hz, n_hzπ, n_z, n_z))
    for a in range(n_hzπ):
        for b in range(n_hz):
            for c in range(n_zπ):
                z[c, b, a], z_Q[c, b, a] = rouwenhorst(n_z, ρ, σ_z[b],
                                                       ρ_π * zπ[a, c])
    return (z, z_Q, zπ, zπ_Q, h_z, Q_hz, σ_z, h_c, Q_hc, σ_c,
            h_zπ, Q_hzπ, σ_zπ, h_λ, Q_hλ)
