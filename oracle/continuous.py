"""Oracle restatement of the continuous-state operators (test infrastructure only).

Follows /root/reference/code/ssy/continuous_junnan/ssy_wc_ratio_continuous.py
(build_grid :20-56, next_state :63-83, Kg_vmap_quad :125-153, Kg_vmap_mc :90-118,
T_fun_factory :156-226), /root/reference/code/gcy/continuous/gcy_wc_ratio_continuous.py
(build_grid :23-70, next_state :77-115) and /root/reference/code/utils.py:6-23 (lin_interp =
map_coordinates(order=1, mode='nearest') on uniform grids).

PARITY UNPINNED: the reference's continuous path needs jax and quantecon.quad.qnwnorm (both
un-vendored, un-installable) and records no numeric output for it.  qnwnorm([d]*dim) is restated
as the tensor product of Gauss-Hermite rules for N(0,1): nodes sqrt(2) x_GH, weights w_GH/sqrt(pi)
(numpy.polynomial.hermite.hermgauss), first dimension varying fastest.
"""
import itertools

import numpy as np
from numpy.polynomial.hermite import hermgauss
from scipy.ndimage import map_coordinates


def qnwnorm(ns):
    xs, ws = [], []
    for n in ns:
        x, w = hermgauss(n)
        xs.append(x * np.sqrt(2.0))
        ws.append(w / np.sqrt(np.pi))
    # first dimension fastest
    idx = list(itertools.product(*[range(n) for n in ns[::-1]]))
    nodes = np.array([[xs[d][t[len(ns) - 1 - d]] for d in range(len(ns))] for t in idx])       # (Q, dim)
    weights = np.array([np.prod([ws[d][t[len(ns) - 1 - d]] for d in range(len(ns))]) for t in idx])
    return nodes, weights


def build_grid_ssy(ssy, sizes, num_std_devs=3.2):
    β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ = ssy.params
    grids = []
    for s, r, n in zip((s_λ, s_c, s_z), (ρ_λ, ρ_c, ρ_z), sizes[:3]):
        g = num_std_devs * np.sqrt(s ** 2 / (1 - r ** 2))
        grids.append(np.linspace(-g, g, n))
    h_z_max = num_std_devs * np.sqrt(s_z ** 2 / (1 - ρ_z ** 2))
    z_max = num_std_devs * ϕ_z * np.exp(h_z_max)
    grids.append(np.linspace(-z_max, z_max, sizes[3]))
    return tuple(grids)


def build_grid_gcy(gcy, sizes, num_std_devs=3.2):
    (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z, ρ_ππ, φ_zπ, ρ_zπ, s_zπ) = gcy.params
    grids = []
    for s, r, n in zip((s_λ, s_c, s_z, s_zπ), (ρ_λ, ρ_c, ρ_z, ρ_zπ), sizes[:4]):
        g = num_std_devs * np.sqrt(s ** 2 / (1 - r ** 2))
        grids.append(np.linspace(-g, g, n))
    h_zπ_max = num_std_devs * np.sqrt(s_zπ ** 2 / (1 - ρ_zπ ** 2))
    σ_zπ_max = φ_zπ * np.exp(h_zπ_max)
    zπ_max = num_std_devs * np.sqrt(σ_zπ_max ** 2 / (1 - ρ_ππ ** 2))
    zπ_grid = np.linspace(-zπ_max, zπ_max, sizes[5])
    h_z_max = num_std_devs * np.sqrt(s_z ** 2 / (1 - ρ_z ** 2))
    σ_z_max = φ_z * np.exp(h_z_max)
    z_max = (ρ_π * zπ_grid[-1] + num_std_devs * σ_z_max) / (1 - ρ)
    z_min = (ρ_π * zπ_grid[0] - num_std_devs * σ_z_max) / (1 - ρ)
    grids.append(np.linspace(z_min, z_max, sizes[4]))
    grids.append(zπ_grid)
    return tuple(grids)


def lin_interp(x, vals, grids):
    """x: (dim, M) points; uniform grids; order-1 interpolation, nearest-edge extension."""
    lo = np.array([g[0] for g in grids])[:, None]
    h = np.array([g[1] - g[0] for g in grids])[:, None]
    return map_coordinates(vals, (x - lo) / h, order=1, mode="nearest")


class ContSSY:
    """State x = (h_lam, h_c, h_z, z); quadrature or Monte-Carlo draws eta of shape (4, Q)."""
    dim = 4

    def __init__(self, ssy, sizes, nodes, weights, num_std_devs=3.2):
        self.params = ssy.params
        self.grids = build_grid_ssy(ssy, sizes, num_std_devs)
        self.shape = tuple(sizes)
        self.nodes = np.asarray(nodes, dtype=np.float64)      # (dim, Q)
        self.weights = np.asarray(weights, dtype=np.float64)  # (Q,)
        β, γ, ψ = self.params[0], self.params[1], self.params[2]
        self.β, self.γ, self.θ = β, γ, (1 - γ) / (1 - 1 / ψ)
        mesh = np.meshgrid(*self.grids, indexing="ij")
        self.X = np.stack([m.ravel() for m in mesh], axis=0)   # (dim, N)

    def next_state(self, x):
        β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ = self.params
        h_λ, h_c, h_z, z = x
        η = self.nodes
        σ_z = ϕ_z * np.exp(h_z)
        return np.array([ρ_λ * h_λ + s_λ * η[0], ρ_c * h_c + s_c * η[1], ρ_z * h_z + s_z * η[2], ρ * z + σ_z * η[3]])

    def const(self, x):
        μ_c, ϕ_c = self.params[3], self.params[6]
        σ_c = ϕ_c * np.exp(x[1])
        return np.exp((1 - self.γ) * (μ_c + x[3]) + 0.5 * (1 - self.γ) ** 2 * σ_c ** 2)

    def _per_state(self, w, fn):
        out = np.empty(self.X.shape[1])
        for n in range(self.X.shape[1]):
            x = self.X[:, n]
            nx = self.next_state(x)
            out[n] = fn(x, nx, np.exp(self.θ * nx[0]), lin_interp(nx, w, self.grids))
        return out.reshape(self.shape)

    def Kg(self, w):
        return self._per_state(w, lambda x, nx, pf, wi: self.const(x) * np.dot(wi ** self.θ * pf, self.weights))

    def T(self, w):
        return 1 + self.β * self.Kg(w) ** (1 / self.θ)

    def jvp(self, w, v):
        θ = self.θ
        s = self.Kg(w)
        ds = self._per_state(w, lambda x, nx, pf, wi: self.const(x) * np.dot(
            θ * wi ** (θ - 1) * lin_interp(nx, v, self.grids) * pf, self.weights))
        return self.β / θ * s ** (1 / θ - 1) * ds


class ContGCY(ContSSY):
    """State x = (h_lam, h_c, h_z, h_zpi, z, z_pi); draws eta of shape (6, Q)."""
    dim = 6

    def __init__(self, gcy, sizes, nodes, weights, num_std_devs=3.2):
        self.params = gcy.params
        self.grids = build_grid_gcy(gcy, sizes, num_std_devs)
        self.shape = tuple(sizes)
        self.nodes = np.asarray(nodes, dtype=np.float64)
        self.weights = np.asarray(weights, dtype=np.float64)
        β, ψ, γ = self.params[0], self.params[1], self.params[2]
        self.β, self.γ, self.θ = β, γ, (1 - γ) / (1 - 1 / ψ)
        mesh = np.meshgrid(*self.grids, indexing="ij")
        self.X = np.stack([m.ravel() for m in mesh], axis=0)

    def next_state(self, x):
        (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z, ρ_ππ, φ_zπ, ρ_zπ, s_zπ) = self.params
        h_λ, h_c, h_z, h_zπ, z, z_π = x
        η = self.nodes
        σ_z = φ_z * np.exp(h_z)
        σ_zπ = φ_zπ * np.exp(h_zπ)
        return np.array([ρ_λ * h_λ + s_λ * η[0], ρ_c * h_c + s_c * η[1], ρ_z * h_z + s_z * η[2],
                         ρ_zπ * h_zπ + s_zπ * η[3], ρ * z + ρ_π * z_π + σ_z * η[4], ρ_ππ * z_π + σ_zπ * η[5]])

    def const(self, x):
        μ_c, φ_c = self.params[5], self.params[6]
        σ_c = φ_c * np.exp(x[1])
        return np.exp((1 - self.γ) * (μ_c + x[4]) + 0.5 * (1 - self.γ) ** 2 * σ_c ** 2)
