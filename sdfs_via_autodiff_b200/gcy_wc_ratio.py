"""GCY wealth-consumption ratio on a discretised Markov grid -- host mirror of
/root/reference/code/gcy/discrete/gcy_wc_ratio.py (discretize_gcy :31-131, T_gcy
:134-238, test_compute_wc_ratio_gcy :319-340).  All arithmetic runs in
libsdfs_b200 on the GPU.
"""
import numpy as np

from .operator import Factors, WCOperator, cached_operator, MODEL_GCY
from .solvers import solver
from .gcy_model import GCY


def discretize_gcy(gcy, shapes, ctx=None):
    """Discretise the GCY model (nested Rouwenhorst chains, drift ρ_π z_π in the z chain)
    on the device and return the reference's 15-tuple of NumPy arrays

        (z, z_Q, z_π, z_π_Q, h_z, h_z_Q, σ_z, h_c, h_c_Q, σ_c, h_zπ, h_zπ_Q, σ_zπ, h_λ, h_λ_Q)."""
    return Factors.build(MODEL_GCY, gcy.params, shapes, ctx).arrays()


def make_T_gcy(gcy_or_params, shapes, arrays=None, storage="auto", ctx=None):
    params = getattr(gcy_or_params, "params", gcy_or_params)
    if arrays is None:
        return WCOperator.from_factors(Factors.build(MODEL_GCY, params, shapes, ctx), storage)
    return WCOperator.from_factors(Factors.from_host(MODEL_GCY, params, shapes, arrays, ctx), storage)


def T_gcy(w, shapes, params, arrays, storage="auto"):
    """Same signature as the reference's jitted ``T_gcy(w, shapes, params, arrays)``."""
    op = cached_operator(MODEL_GCY, shapes, params, arrays, storage)
    return op(w)


def T_gcy_loops(w, shapes, params, arrays):
    """Counterpart of the reference's loop form (gcy_wc_ratio.py:244-302): T from the explicit single-index
    transition matrix (dense-storage operator), independent of the sum-factorised ``T_gcy``."""
    return cached_operator(MODEL_GCY, shapes, params, arrays, "dense")(w)


def test_vectorized_equals_loops(shapes=(2, 3, 4, 5, 6, 7)):
    """gcy_wc_ratio.py:305-316: the factor-form T and the explicit-matrix T agree at a random w."""
    gcy = GCY()
    params = gcy.params
    arrays = discretize_gcy(gcy, shapes)
    w = np.exp(np.random.randn(*shapes))  # Test operator at w
    w1 = T_gcy(w, shapes, params, arrays, storage="kron")
    w2 = T_gcy_loops(w, shapes, params, arrays)
    same = bool(np.allclose(np.asarray(w1), np.asarray(w2)))
    print(same)
    return same


test_vectorized_equals_loops.__test__ = False   # a driver, not a pytest test


def test_compute_wc_ratio_gcy(shapes=(3, 3, 3, 3, 3, 3), algo="successive_approx"):
    """Solve a small version of the model using T_gcy."""
    gcy = GCY()
    params = gcy.params
    arrays = discretize_gcy(gcy, shapes)
    T = lambda w: T_gcy(w, shapes, params, arrays)
    init_val = 800.0
    w_init = np.ones(shapes) * init_val
    w_star = solver(T, w_init, algorithm=algo)
    return w_star


test_compute_wc_ratio_gcy.__test__ = False
