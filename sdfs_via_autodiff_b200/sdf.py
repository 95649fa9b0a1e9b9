"""Solve functions that return w* AND the SDF (north-star addition; the reference has
no SDF code -- formulas from paper/autosdfs.tex:374-384)."""
from .solvers import newton_solver, successive_approx
from .ssy_wc_ratio import make_T_ssy
from .gcy_wc_ratio import make_T_gcy


class SDFResult:
    """w: fixed point; q_f: one-period risk-free price E[M'|x]; euler: Euler-equation
    residual (≈0 at the fixed point); op: operator (``op.sdf_rows(w, rows)`` gives M̄ rows)."""

    def __init__(self, op, w, q_f, euler, iters, info):
        self.op, self.w, self.q_f, self.euler, self.iters, self.info = op, w, q_f, euler, iters, info

    def sdf_rows(self, rows):
        return self.op.sdf_rows(self.w, rows)


def _solve(op, algo, init_val, verbose, **kw):
    w0 = op.ctx.full(op.shapes, init_val)
    if algo == "newton":
        w, k, info = newton_solver(op, w0, verbose=verbose, return_info=True, **kw)
    else:
        w, k, info = successive_approx(op, w0, verbose=verbose, return_info=True, **kw)
    q_f, euler = op.sdf(w)
    return SDFResult(op, w, q_f, euler, k, info)


def solve_ssy(ssy, shapes, algo="newton", storage="auto", init_val=800.0, verbose=False, **kw):
    return _solve(make_T_ssy(ssy, shapes, storage=storage), algo, init_val, verbose, **kw)


def solve_gcy(gcy, shapes, algo="newton", storage="auto", init_val=800.0, verbose=False, **kw):
    return _solve(make_T_gcy(gcy, shapes, storage=storage), algo, init_val, verbose, **kw)
