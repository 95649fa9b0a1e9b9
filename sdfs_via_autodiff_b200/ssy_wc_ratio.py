"""SSY wealth-consumption ratio on a discretised Markov grid -- host mirror of
/root/reference/code/ssy/discrete/ssy_wc_ratio.py (discretize_ssy :23-79, T_ssy
:82-151, test_compute_wc_ratio_ssy :216-240).  All arithmetic runs in
libsdfs_b200 on the GPU.
"""
import time

import numpy as np

from .operator import Factors, WCOperator, cached_operator, MODEL_SSY
from .solvers import solver
from .ssy_model import SSY


def discretize_ssy(ssy, shapes, ctx=None):
    """Discretise the SSY model with iterated Rouwenhorst chains (computed on the device)
    and return the reference's 10-tuple of NumPy arrays

        (h_λ, h_λ_Q, h_c, h_c_Q, h_z, h_z_Q, z[i, j], z_Q[i, j, jp], σ_c, σ_z)."""
    return Factors.build(MODEL_SSY, ssy.params, shapes, ctx).arrays()


def make_T_ssy(ssy_or_params, shapes, arrays=None, storage="auto", ctx=None):
    """Operator handle for the SSY model.  ``arrays=None`` discretises on the device."""
    params = getattr(ssy_or_params, "params", ssy_or_params)
    if arrays is None:
        return WCOperator.from_factors(Factors.build(MODEL_SSY, params, shapes, ctx), storage)
    return WCOperator.from_factors(Factors.from_host(MODEL_SSY, params, shapes, arrays, ctx), storage)


def T_ssy(w, shapes, params, arrays, storage="auto"):
    """Discrete operator T for the SSY model: same signature as the reference's jitted
    ``T_ssy(w, shapes, params, arrays)``; returns a DeviceArray of shape ``shapes``."""
    op = cached_operator(MODEL_SSY, shapes, params, arrays, storage)
    return op(w)


def T_ssy_loops(w, shapes, params, arrays):
    """Counterpart of the reference's loop form (ssy_wc_ratio.py:159-199): T evaluated from the explicit
    single-index transition matrix P (one row-times-vector reduction per state, temp_ssy.py:106), i.e. the
    dense-storage operator - an implementation independent of the sum-factorised one behind ``T_ssy``."""
    return cached_operator(MODEL_SSY, shapes, params, arrays, "dense")(w)


def test_vectorized_equals_loops(shapes=(4, 7, 6, 5)):
    """ssy_wc_ratio.py:202-213: the factor-form T and the explicit-matrix T agree at a random w."""
    ssy = SSY()
    params = ssy.params
    arrays = discretize_ssy(ssy, shapes)
    w = np.exp(np.random.randn(*shapes))  # Test operator at w
    w1 = T_ssy(w, shapes, params, arrays, storage="kron")
    w2 = T_ssy_loops(w, shapes, params, arrays)
    same = bool(np.allclose(np.asarray(w1), np.asarray(w2)))
    print(same)
    return same


test_vectorized_equals_loops.__test__ = False   # a driver, not a pytest test


def test_compute_wc_ratio_ssy(shapes=(2, 3, 4, 5), algo="successive_approx"):
    """Solve a small version of the model using T_ssy."""
    ssy = SSY()
    params = ssy.params
    arrays = discretize_ssy(ssy, shapes)
    T = lambda w: T_ssy(w, shapes, params, arrays)
    init_val = 800.0
    w_init = np.ones(shapes) * init_val
    t0 = time.time()
    w_star = solver(T, w_init, algorithm=algo)
    t = time.time() - t0
    print(f"Computed solution in {t} seconds.")
    return w_star


test_compute_wc_ratio_ssy.__test__ = False      # a driver, not a pytest test
