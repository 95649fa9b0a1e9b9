"""Fixed-point solvers -- host mirror of /root/reference/code/solvers.py.

Same names, keyword arguments, return values and printed lines as the reference
(successive_approx :19-48, newton_solver :51-95, solvers dict :146-149, solver
:154-177).  When ``f`` is (or wraps) a device operator the whole loop runs inside
one cooperative CUDA kernel with no host synchronisation per iteration; the
"iter = k, error = e" lines are reconstructed from the device-side history.
"""
import ctypes as C
from textwrap import dedent

from ._lib import lib, check
from .operator import resolve_operator

default_tolerance = 1e-7
default_max_iter = int(1e6)

KRYLOV = {"bicgstab": 0, "gmres": 1}

_NOT_AN_OPERATOR = ("{} needs an operator of this package (WCOperator, or a closure over T_ssy/T_gcy such as "
                    "`lambda w: T_ssy(w, shapes, params, arrays)`): the loop runs inside a CUDA kernel; there is "
                    "no host-side iteration of arbitrary Python callables and no CPU fallback")


def _report(history, current_iter, max_iter, verbose, print_skip, stride=1):
    """Print what the reference's loop prints (solvers.py:37-46) from the device-side history:
    history[i // stride] is the error of iteration i for i % stride == 0."""
    if verbose and history is not None:
        for i in range(0, current_iter, print_skip):
            if i % stride == 0 and i // stride < len(history):
                print("iter = {}, error = {}".format(i, history[i // stride]))
    _finish_messages(current_iter, max_iter, verbose)


def _finish_messages(current_iter, max_iter, verbose):
    if current_iter == max_iter:
        print(f"Warning: Hit maximum iteration number {max_iter}")
    elif verbose:
        print(f"Iteration converged after {current_iter} iterations")


def successive_approx(f, x_init, tol=default_tolerance, max_iter=default_max_iter,
                      verbose=True, print_skip=1000, return_info=False):
    "Uses successive approximation on f."
    op = resolve_operator(f)
    if op is None:
        raise TypeError(_NOT_AN_OPERATOR.format("successive_approx"))
    max_iter = int(max_iter)
    if verbose:
        print("Beginning iteration\n\n")
    ctx = op.ctx
    w0 = op._in(x_init)
    w_out = ctx.empty(op.shapes)
    iters, ferr = C.c_int64(), C.c_double()
    hist = None
    cap = 0
    if verbose:
        cap = min(max_iter // max(1, print_skip) + 1, 1 << 20)
        hist = ctx.empty((cap,))
    check(lib.sdfs_solve_sa(op.handle, w0.ptr, float(tol), max_iter, w_out.ptr, C.byref(iters), C.byref(ferr),
                            hist.ptr if hist is not None else None, int(max(1, print_skip)), cap), ctx.handle)
    k = iters.value
    _report(hist.numpy() if verbose else None, k, max_iter, verbose, print_skip, stride=max(1, print_skip))
    if return_info:
        return w_out, k, dict(final_error=ferr.value)
    return w_out, k


def newton_solver(f, x_init, tol=default_tolerance, max_iter=default_max_iter,
                  bicgstab_atol=1e-4, verbose=True, print_skip=1,
                  krylov="bicgstab", krylov_rtol=1e-5, restart=30, krylov_maxiter=None,
                  return_info=False):
    """Newton's method on g(x) = f(x) - x:  x <- x - J_g(x)^{-1} g(x), iterated with the
    successive-approximation stopping rule (solvers.py:83-95).

    J_g(x) v is the analytic Jacobian-vector product of T (replaces jax.jvp); the
    linear solve is an on-device BiCGSTAB with the recurrence and stopping rule of
    ``jax.scipy.sparse.linalg.bicgstab(..., atol=bicgstab_atol)`` (``krylov="bicgstab"``,
    the reference's choice) or restarted GMRES (``krylov="gmres"``)."""
    op = resolve_operator(f)
    if op is None:
        raise TypeError(_NOT_AN_OPERATOR.format("newton_solver"))
    max_iter = int(max_iter)
    if verbose:
        print("Beginning iteration\n\n")
    ctx = op.ctx
    w0 = op._in(x_init)
    w_out = ctx.empty(op.shapes)
    cap = 120
    outer, ferr, nmv = C.c_int64(), C.c_double(), C.c_int64()
    h_err = (C.c_double * cap)()
    h_inner = (C.c_int64 * cap)()
    check(lib.sdfs_solve_newton(op.handle, w0.ptr, float(tol), max_iter, KRYLOV[krylov], float(krylov_rtol),
                                float(bicgstab_atol), int(restart),
                                int(krylov_maxiter) if krylov_maxiter else 0, w_out.ptr, C.byref(outer),
                                C.byref(ferr), h_err, h_inner, cap, C.byref(nmv)), ctx.handle)
    k = outer.value
    _report(list(h_err[:min(k, cap)]) if verbose else None, k, max_iter, verbose, print_skip, stride=1)
    if verbose and k > cap:
        print(f"(error history is recorded for the first {cap} outer iterations only)")
    if return_info:
        n = min(k, cap)
        return w_out, k, dict(final_error=ferr.value, errors=list(h_err[:n]),
                              inner_iters=list(h_inner[:n]), matvecs=nmv.value)
    return w_out, k


def anderson_solver(f, x_init, tol=default_tolerance, max_iter=10000, verbose=True, return_info=False):
    """Anderson acceleration with the reference's hard-coded parameters (solvers.py:98-124:
    history_size=10, mixing_frequency=4, beta=8.0, ridge=1e-6), device resident.  jaxopt is not
    vendored with the reference, so its update rule is restated (see include/sdfs_b200.h):
    results agree with the package's oracle, parity with jaxopt itself is unpinned."""
    op = resolve_operator(f)
    if op is None:
        raise TypeError(_NOT_AN_OPERATOR.format("anderson_solver"))
    max_iter = int(max_iter)
    ctx = op.ctx
    w0 = op._in(x_init)
    w_out = ctx.empty(op.shapes)
    iters, ferr = C.c_int64(), C.c_double()
    check(lib.sdfs_solve_anderson(op.handle, w0.ptr, float(tol), max_iter, 10, 4, 8.0, 1e-6, w_out.ptr,
                                  C.byref(iters), C.byref(ferr)), ctx.handle)
    current_iter = iters.value
    _finish_messages(current_iter, max_iter, verbose)
    if return_info:
        return w_out, current_iter, dict(final_error=ferr.value)
    return w_out, current_iter


def fixed_point_via_gradient_decent(f, x_init):
    raise NotImplementedError("gradient descent (jaxopt wrapper, solvers.py:127-140) is outside the "
                              "accelerated hot path; use 'newton' or 'successive_approx'")


# A dictionary of available solvers (same keys as the reference).
solvers = dict((("newton", newton_solver),
                ("anderson", anderson_solver),
                ("gd", fixed_point_via_gradient_decent),
                ("successive_approx", successive_approx)))


def solver(f, x_init, algorithm="newton", verbose=True):
    """A simple front end to the other solvers (defaults only, returns x* only)."""
    try:
        solver = solvers[algorithm]
    except KeyError:
        msg = f"""\
                  Algorithm {algorithm} not found.  
                  Falling back to successive approximation.
               """
        print(dedent(msg))
        solver = successive_approx
    # the reference calls the chosen solver with its defaults only (solvers.py:175: tol and verbose
    # are not forwarded, so it always prints); verbose is honoured here so callers can silence it
    if verbose:
        x_star, num_iter = solver(f, x_init)
    else:
        x_star, num_iter = solver(f, x_init, verbose=False)
    return x_star
