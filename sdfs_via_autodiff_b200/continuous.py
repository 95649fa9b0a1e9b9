"""Continuous-state wealth-consumption operators -- host mirror of
/root/reference/code/ssy/continuous_junnan/ssy_wc_ratio_continuous.py and
/root/reference/code/gcy/continuous/gcy_wc_ratio_continuous.py (build_grid, T_fun_factory,
wc_ratio_continuous).  The conditional expectation uses Gauss-Hermite quadrature (the
reference's quantecon.quad.qnwnorm([d]*dim), restated with numpy's hermgauss) or Monte-Carlo
draws; w is interpolated multilinearly on uniform grids (utils.py:6-23).  All arithmetic runs in
libsdfs_b200 (csrc/cont.cuh); the solvers are the same device-resident loops as for the
discretised models.  Parity with the JAX original is unpinned (jax/quantecon not installable,
no recorded outputs); the package oracle restates it independently (oracle/continuous.py).
"""
import ctypes as C
import itertools

import numpy as np
from numpy.polynomial.hermite import hermgauss

from ._lib import lib, check
from .device import Context
from .operator import WCOperator, MODEL_SSY, MODEL_GCY
from .solvers import solvers as _solver_table, successive_approx as _successive_approx, solver as _solver


def _is_gcy(model):
    return hasattr(model, "ρ_ππ")


def build_grid(model, *sizes, num_std_devs=3.2):
    """Interpolation grids: SSY (h_λ, h_c, h_z, z) sizes, GCY (h_λ, h_c, h_z, h_zπ, z, z_π) sizes."""
    if len(sizes) == 1 and hasattr(sizes[0], "__len__"):
        sizes = tuple(sizes[0])
    if _is_gcy(model):
        (β, ψ, γ, ρ_λ, s_λ, μ_c, φ_c, ρ, ρ_π, φ_z, ρ_c, s_c, ρ_z, s_z, ρ_ππ, φ_zπ, ρ_zπ, s_zπ) = model.params
        grids = []
        for s, r, n in zip((s_λ, s_c, s_z, s_zπ), (ρ_λ, ρ_c, ρ_z, ρ_zπ), sizes[:4]):
            g_max = num_std_devs * np.sqrt(s ** 2 / (1 - r ** 2))
            grids.append(np.linspace(-g_max, g_max, n))
        h_zπ_max = num_std_devs * np.sqrt(s_zπ ** 2 / (1 - ρ_zπ ** 2))
        σ_zπ_max = φ_zπ * np.exp(h_zπ_max)
        zπ_max = num_std_devs * np.sqrt(σ_zπ_max ** 2 / (1 - ρ_ππ ** 2))
        zπ_grid = np.linspace(-zπ_max, zπ_max, sizes[5])
        h_z_max = num_std_devs * np.sqrt(s_z ** 2 / (1 - ρ_z ** 2))
        σ_z_max = φ_z * np.exp(h_z_max)
        z_max = (ρ_π * zπ_grid[-1] + num_std_devs * σ_z_max) / (1 - ρ)
        z_min = (ρ_π * zπ_grid[0] - num_std_devs * σ_z_max) / (1 - ρ)
        grids += [np.linspace(z_min, z_max, sizes[4]), zπ_grid]
        return tuple(grids)
    β, γ, ψ, μ_c, ρ, ϕ_z, ϕ_c, ρ_z, ρ_c, ρ_λ, s_z, s_c, s_λ = model.params
    grids = []
    for s, r, n in zip((s_λ, s_c, s_z), (ρ_λ, ρ_c, ρ_z), sizes[:3]):
        g_max = num_std_devs * np.sqrt(s ** 2 / (1 - r ** 2))
        grids.append(np.linspace(-g_max, g_max, n))
    h_z_max = num_std_devs * np.sqrt(s_z ** 2 / (1 - ρ_z ** 2))
    z_max = num_std_devs * ϕ_z * np.exp(h_z_max)
    grids.append(np.linspace(-z_max, z_max, sizes[3]))
    return tuple(grids)


def gauss_hermite_normal(d, dim):
    """Nodes (dim, d**dim) and weights of the tensor-product Gauss-Hermite rule for N(0, I)
    (what quantecon.quad.qnwnorm([d]*dim) returns, first dimension varying fastest)."""
    x, w = hermgauss(d)
    x, w = x * np.sqrt(2.0), w / np.sqrt(np.pi)
    idx = np.array(list(itertools.product(range(d), repeat=dim)))[:, ::-1]      # first dimension fastest
    return np.ascontiguousarray(x[idx].T), np.prod(w[idx], axis=1)


def T_fun_factory(params, method="quadrature", batch_size=None, model_kind=None, ctx=None):
    """Operator T for the continuous-state model.  ``params`` as in the reference:
    (model_params, grids, nodes, weights) for "quadrature", (model_params, grids, mc_draws) for
    "monte_carlo" (nodes / mc_draws of shape (dim, Q)).  ``batch_size`` is accepted and ignored:
    the kernel needs no batching."""
    ctx = ctx or Context.default()
    if method == "quadrature":
        model_params, grids, nodes, weights = params
    elif method == "monte_carlo":
        model_params, grids, nodes = params
        nodes = np.asarray(nodes, dtype=np.float64)
        weights = np.full(nodes.shape[1], 1.0 / nodes.shape[1])               # jnp.mean
    else:
        raise KeyError("Method not found.")
    grids = [np.ascontiguousarray(np.asarray(g, dtype=np.float64)) for g in grids]
    dim = len(grids)
    kind = model_kind if model_kind is not None else (MODEL_SSY if dim == 4 else MODEL_GCY)
    nodes = np.ascontiguousarray(np.asarray(nodes, dtype=np.float64))
    weights = np.ascontiguousarray(np.asarray(weights, dtype=np.float64))
    if nodes.shape != (dim, weights.size):
        raise ValueError(f"nodes must have shape ({dim}, Q), got {nodes.shape}")
    sizes = (C.c_int32 * dim)(*[g.size for g in grids])
    flat = np.ascontiguousarray(np.concatenate(grids))
    p = (C.c_double * len(model_params))(*[float(v) for v in model_params])
    h = C.c_void_p()
    dp = C.POINTER(C.c_double)
    check(lib.sdfs_op_continuous(ctx.handle, kind, p, sizes, flat.ctypes.data_as(dp), nodes.ctypes.data_as(dp),
                                 weights.ctypes.data_as(dp), weights.size, C.byref(h)), ctx.handle)
    return WCOperator(ctx, h, tuple(g.size for g in grids))


def make_T_continuous(model, sizes, method="quadrature", d=5, mc_draws=None, mc_draw_size=2000, seed=1234,
                      num_std_devs=3.2, ctx=None):
    grids = build_grid(model, *sizes, num_std_devs=num_std_devs)
    dim = len(grids)
    if method == "quadrature":
        nodes, weights = gauss_hermite_normal(d, dim)
        params = (model.params, grids, nodes, weights)
    elif method == "monte_carlo":
        if mc_draws is None:      # the reference draws with jax.random.PRNGKey(seed); NumPy's generator here
            mc_draws = np.random.default_rng(seed).standard_normal((dim, mc_draw_size))
        params = (model.params, grids, mc_draws)
    else:
        raise KeyError("Approximation method not found.")
    return grids, T_fun_factory(params, method, model_kind=MODEL_GCY if _is_gcy(model) else MODEL_SSY, ctx=ctx)


_GRID_KEYS = {False: ("h_λ_grid_size", "h_c_grid_size", "h_z_grid_size", "z_grid_size"),
              True: ("h_λ_grid_size", "h_c_grid_size", "h_z_grid_size", "h_zπ_grid_size", "z_grid_size",
                     "z_π_grid_size")}
_GRID_DEFAULTS = {False: (10, 10, 10, 20), True: (10, 10, 10, 10, 20, 20)}


def wc_ratio_continuous(model, *grid_sizes, num_std_devs=3.2, d=5, mc_draw_size=2000, seed=1234, w_init=None,
                        ram_free=20, tol=1e-5, method="quadrature", algorithm="successive_approx", verbose=True,
                        write_to_file=True, filename="w_star_data.npy", mc_draws=None, solver_tol=None,
                        **grid_kw):
    """Iterate to convergence on the continuous-state operator and return (grids, w_star).

    Same signature and defaults as the reference (ssy_wc_ratio_continuous.py:229-297,
    gcy_wc_ratio_continuous.py:264-335): the grid sizes may be given positionally or with the
    reference's keyword names (``h_λ_grid_size=...``, ``z_grid_size=...``; GCY adds
    ``h_zπ_grid_size`` and ``z_π_grid_size``), ``w_init`` defaults to ones, ``write_to_file``
    defaults to True (two consecutive np.save records).  Like the reference, the solve goes
    through ``solver(T, w_init, algorithm=algorithm)`` -- the reference accepts ``tol`` but never
    forwards it, so the effective tolerance is solvers.py's default 1e-7; ``tol`` is accepted and
    ignored here for the same result.  Extensions: ``solver_tol`` (forwarded to the solver when
    given), ``verbose`` is honoured, ``mc_draws`` supplies the Monte-Carlo shocks; ``ram_free``
    is accepted for compatibility (no batching is needed on the device)."""
    gcy = _is_gcy(model)
    keys = _GRID_KEYS[gcy]
    sizes = list(_GRID_DEFAULTS[gcy])
    if len(grid_sizes) > len(keys):
        raise TypeError(f"wc_ratio_continuous takes at most {len(keys)} grid sizes for this model")
    for i, v in enumerate(grid_sizes):
        sizes[i] = v
    for k, v in grid_kw.items():
        if k not in keys:
            raise TypeError(f"wc_ratio_continuous() got an unexpected keyword argument {k!r}")
        if keys.index(k) < len(grid_sizes):
            raise TypeError(f"wc_ratio_continuous() got multiple values for argument {k!r}")
        sizes[keys.index(k)] = v
    grids, T = make_T_continuous(model, tuple(int(v) for v in sizes), method, d, mc_draws, mc_draw_size, seed,
                                 num_std_devs)
    if w_init is None:
        w_init = T.ctx.full(T.shapes, 1.0)
    if solver_tol is None:
        w_star = _solver(T, w_init, algorithm=algorithm, verbose=verbose)
    else:
        try:
            fn = _solver_table[algorithm]
        except KeyError:
            print(f"Algorithm {algorithm} not found.  \nFalling back to successive approximation.\n")
            fn = _successive_approx
        w_star, _ = fn(T, w_init, tol=solver_tol, verbose=verbose)
    if write_to_file:
        save_wstar(filename, grids, w_star)
    return grids, w_star


def compare_T_factories(T_fact_old, T_fact_new, shape=(5, 6, 7, 8), seed=1234, n=100):
    """Compare the results and speed of two function factories for T (ssy_wc_ratio_continuous.py:330-452):
    both are called as ``fact(params_quad, 'quadrature', batch_size)`` on the reference's test set-up (SSY,
    3 standard deviations, 4 quadrature nodes per dimension), applied to ``n`` random w, and one Newton step
    of each is compared as well.  Prints the reference's report lines and returns True when both agree."""
    import time
    from .ssy_model import SSY
    from .solvers import newton_solver
    ssy = SSY()
    grids = build_grid(ssy, *shape, num_std_devs=3.0)
    nodes, weights = gauss_hermite_normal(4, len(grids))
    params_quad = (ssy.params, grids, nodes, weights)
    batch_size = int(np.prod(shape))
    T_old = T_fact_old(params_quad, "quadrature", batch_size)
    T_new = T_fact_new(params_quad, "quadrature", batch_size)
    ctx = T_new.ctx if hasattr(T_new, "ctx") else Context.default()
    print("----- Testing the Operator T -----")
    w0 = np.zeros(shape) + 1.0
    t0 = time.time(); T_old(w0); ctx.sync(); t1 = time.time(); T_new(w0); ctx.sync(); t2 = time.time()
    print("Compilation time: {:.4f}ms vs {:.4f}ms".format((t1 - t0) * 1000, (t2 - t1) * 1000))
    w0_array = np.random.default_rng(seed).uniform(size=(n,) + tuple(shape)) + 0.5
    t0 = time.time()
    old = [np.asarray(T_old(w0_array[i])) for i in range(n)]
    t1 = time.time()
    new = [np.asarray(T_new(w0_array[i])) for i in range(n)]
    t2 = time.time()
    same_T = all(np.allclose(a, b) for a, b in zip(old, new))
    print("Speed comparison for {} runs: {:.4f}ms vs {:.4f}ms".format(n, 1000 * (t1 - t0), 1000 * (t2 - t1)))
    print("Same results? {}".format(same_T))
    print("\n----- Testing Newton's Method -----")
    m = max(1, int(n / 50))
    t0 = time.time()
    old = [np.asarray(newton_solver(T_old, 400.0 * w0_array[i], max_iter=1, verbose=False)[0]) for i in range(m)]
    t1 = time.time()
    new = [np.asarray(newton_solver(T_new, 400.0 * w0_array[i], max_iter=1, verbose=False)[0]) for i in range(m)]
    t2 = time.time()
    print("Speed comparison for {} runs: {:.4f}s vs {:.4f}s".format(m, t1 - t0, t2 - t1))
    same_N = all(np.allclose(a, b) for a, b in zip(old, new))
    print("Same results? {}".format(same_N))
    return same_T and same_N


def save_wstar(filename, grids, w_star):
    """The reference's on-disk format (ssy_wc_ratio_continuous.py:291-295): two consecutive
    np.save records in one file, the grids then w_star."""
    same = len({len(g) for g in grids}) == 1
    with open(filename, "wb") as f:
        np.save(f, np.asarray(grids) if same else np.array([np.asarray(g) for g in grids], dtype=object),
                allow_pickle=not same)
        np.save(f, np.asarray(w_star))


def lin_interp(x, fun_vals, grids, ctx=None):
    """utils.py:17-23 on the device: x of shape (dim, M) -> M interpolated values."""
    ctx = ctx or Context.default()
    grids = [np.asarray(g, dtype=np.float64) for g in grids]
    dim = len(grids)
    xd = ctx.asarray(np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(dim, -1))) \
        if not hasattr(x, "ptr") else x
    vals = ctx.asarray(fun_vals)
    M = xd.size // dim
    out = ctx.empty((M,))
    sizes = (C.c_int32 * dim)(*[g.size for g in grids])
    g0 = (C.c_double * dim)(*[float(g[0]) for g in grids])
    intv = (C.c_double * dim)(*[float(g[1] - g[0]) for g in grids])
    check(lib.sdfs_interp_points(ctx.handle, dim, sizes, g0, intv, vals.ptr, xd.ptr, M, out.ptr), ctx.handle)
    return out


def construct_wstar_callable(w_star_vals=None, grids=None, datafile="w_star_data.npy"):
    """Callable x -> w*(x) by linear interpolation over the grid (ssy_wc_ratio_continuous.py:304-326);
    data are read from ``datafile`` when not given.  The values stay resident on the device."""
    if w_star_vals is None or grids is None:
        with open(datafile, "rb") as f:
            grids = np.load(f, allow_pickle=True)
            w_star_vals = np.load(f)
    grids = [np.asarray(g, dtype=np.float64) for g in grids]
    vals = Context.default().asarray(w_star_vals)

    def w_star_func(x):
        return lin_interp(x, vals, grids)
    return w_star_func
