"""Batched (γ, ψ, β) parameter sweeps (BASELINE config 5).

All parameter sets share one transition matrix P (P depends only on the ρ's and s's),
so one T step for B parameter sets is S = P·V with per-column prologue/epilogue
(csrc/sweep.cu): either the fp64 tensor-core GEMM against the stored P (``form="dense"``, the
form BASELINE config 5 names) or the sum-factorised contraction over the Markov factors batched
over the columns (``form="factor"``: 16·N·B bytes per mode instead of 2·N²·B flop, no P in
memory).  Across GPUs the columns are sharded; no collective is needed until the results are
gathered.
"""
import ctypes as C

import numpy as np

from ._lib import lib, check
from .device import Context
from .operator import Factors, WCOperator, MODEL_SSY, MODEL_GCY, STORAGE_KRON

STORAGE_DENSE_REPLICATED = 2
STORAGE_KRON_LOCAL = 4          # factor form, whole on every rank (columns, not slabs, are sharded)
SWEEP_DENSE, SWEEP_FACTOR = 0, 1


def make_sweep_operator(model, shapes, ctx=None, form="factor"):
    """Operator for batched sweeps on this rank (columns, not rows, are sharded).

    ``form="factor"`` (default): nothing but the Markov factors is stored; every step contracts
    them mode by mode for all columns at once (any grid size; 12x faster than the GEMM at
    N = 10^4, B = 4096).
    ``form="dense"``: the full P is stored and every step is one fp64 tensor-core GEMM (the form
    BASELINE config 5 names; same results to rounding, same per-column iteration counts)."""
    if form not in ("dense", "factor"):
        raise ValueError("form must be 'dense' or 'factor'")
    ctx = ctx or Context.default()
    kind = MODEL_GCY if hasattr(model, "ρ_ππ") else MODEL_SSY
    fac = Factors.build(kind, model.params, shapes, ctx)
    h = C.c_void_p()
    storage = STORAGE_DENSE_REPLICATED if form == "dense" else STORAGE_KRON_LOCAL
    check(lib.sdfs_op_from_factors(ctx.handle, fac.handle, storage, C.byref(h)), ctx.handle)
    op = WCOperator(ctx, h, shapes, keep=[fac])
    check(lib.sdfs_sweep_set_form(h, SWEEP_DENSE if form == "dense" else SWEEP_FACTOR), ctx.handle)
    return op


def column_slice(B, nranks, rank):
    chunk = (B + nranks - 1) // nranks
    b = min(B, chunk * rank)
    return b, min(B, b + chunk)


def _prefs(prefs):
    p = np.ascontiguousarray(np.asarray(prefs, dtype=np.float64))
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError("prefs must have shape (B, 3): columns (γ, ψ, β)")
    return p


def sweep_apply_T(op, prefs, W):
    """One batched T step.  W: (B, *shapes) host or device array; returns a DeviceArray."""
    p = _prefs(prefs)
    B = p.shape[0]
    d = op.ctx.asarray(W)
    if d.size != B * op.N:
        raise ValueError(f"W has {d.size} elements, expected {B} x {op.N}")
    out = op.ctx.empty((B,) + op.shapes)
    check(lib.sdfs_sweep_apply_T(op.handle, p.ctypes.data_as(C.POINTER(C.c_double)), B, d.ptr, out.ptr),
          op.ctx.handle)
    return out


def sweep_solve(op, prefs, w_init=800.0, tol=1e-7, max_iter=int(1e6), exchange=None, algorithm="successive_approx",
                bicgstab_atol=1e-4, krylov_rtol=1e-5, krylov_maxiter=None, return_info=False):
    """Solve every parameter set at once.  ``algorithm="successive_approx"``: each column follows
    the reference's stopping rule independently and is frozen on the device once converged.
    ``algorithm="newton"``: each column runs the reference's Newton iteration with its own
    BiCGSTAB; all columns share every Krylov mat-vec as one fp64 tensor-core GEMM.

    Returns (W, iters, final_err): W is a DeviceArray (B_local, *shapes).  With ``exchange``
    (see dist.TorchExchange) the B columns are split over the ranks of ``op.ctx`` and the
    per-rank results are gathered on the host: returns NumPy arrays for all B columns."""
    p = _prefs(prefs)
    ctx = op.ctx
    B = p.shape[0]
    b0, b1 = (0, B) if exchange is None else column_slice(B, ctx.nranks, ctx.rank)
    loc = np.ascontiguousarray(p[b0:b1])
    nb = b1 - b0
    out = ctx.empty((max(nb, 1),) + op.shapes)
    iters = (C.c_int64 * max(nb, 1))()
    errs = (C.c_double * max(nb, 1))()
    info = {}
    if nb > 0 and algorithm == "newton":
        inner = (C.c_int64 * nb)()
        gemms = C.c_int64()
        check(lib.sdfs_sweep_solve_newton(op.handle, loc.ctypes.data_as(C.POINTER(C.c_double)), nb, float(w_init),
                                          float(tol), int(max_iter), float(krylov_rtol), float(bicgstab_atol),
                                          int(krylov_maxiter) if krylov_maxiter else 0, out.ptr, iters, errs, inner,
                                          C.byref(gemms)), ctx.handle)
        info = dict(inner_total=np.array(inner[:nb], dtype=np.int64), gemms=gemms.value)
    elif nb > 0:
        if algorithm != "successive_approx":
            raise KeyError(algorithm)
        check(lib.sdfs_sweep_solve_sa(op.handle, loc.ctypes.data_as(C.POINTER(C.c_double)), nb, float(w_init),
                                      float(tol), int(max_iter), out.ptr, iters, errs), ctx.handle)
    it = np.array(iters[:nb], dtype=np.int64)
    er = np.array(errs[:nb], dtype=np.float64)
    if exchange is None:
        return (out, it, er, info) if return_info else (out, it, er)
    if not hasattr(exchange, "allgather"):
        from .dist import TorchExchange
        exchange = TorchExchange(exchange)
    parts = exchange.allgather((np.asarray(out)[:nb], it, er))
    return (np.concatenate([q[0] for q in parts]), np.concatenate([q[1] for q in parts]),
            np.concatenate([q[2] for q in parts]))
