"""Oracle fixed-point solvers (test infrastructure only).

successive_approx / newton_solver / solver follow
/root/reference/code/solvers.py:19-48, 51-95, 146-177.

The Newton inner solve in the reference is
``jax.scipy.sparse.linalg.bicgstab(jac_x_prod, g(x), atol=1e-4)`` (solvers.py:91).
JAX (< 0.4.25 per the ``jax.config`` import at solvers.py:9) is an un-vendored,
un-pinned third-party dependency and not installable here, so its published
BiCGSTAB recurrence (jax/_src/scipy/sparse/linalg.py ``_bicgstab_solve``) is
restated in ``bicgstab_jax``: x0 = 0, M = I, tol = 1e-5, maxiter = 10*size,
atol2 = max(tol^2 <b,b>, atol^2), loop while <r,r> > atol2 and 0 <= k < maxiter,
early-exit half step, breakdown codes -10/-11.  jax.jvp through T is replaced by
the exact analytic derivative (temp_ssy.py:204-216).
Pin: the recorded Newton trace in sandpit.ipynb (3-7 digits).
"""
import numpy as np

default_tolerance = 1e-7
default_max_iter = int(1e6)


def successive_approx(f, x_init, tol=default_tolerance, max_iter=default_max_iter,
                      verbose=True, print_skip=1000, history=None):
    if verbose:
        print("Beginning iteration\n\n")
    k = 0
    x = x_init
    error = tol + 1
    while error > tol and k < max_iter:
        x_new = f(x)
        error = np.max(np.abs(x_new - x))
        if history is not None:
            history.append(float(error))
        if verbose and k % print_skip == 0:
            print("iter = {}, error = {}".format(k, error))
        k += 1
        x = x_new
    if k == max_iter:
        print(f"Warning: Hit maximum iteration number {max_iter}")
    elif verbose:
        print(f"Iteration converged after {k} iterations")
    return x, k


def bicgstab_jax(A, b, tol=1e-5, atol=0.0, maxiter=None, info=None):
    """BiCGSTAB with the recurrence and stopping rule of JAX's implementation."""
    b = np.asarray(b, dtype=np.float64)
    if maxiter is None:
        maxiter = 10 * b.size
    x = np.zeros_like(b)
    bs = float(np.vdot(b, b))
    atol2 = max(tol ** 2 * bs, atol ** 2)
    r = b - A(x)
    rhat = r.copy()
    alpha = omega = rho = 1.0
    p = r.copy()
    q = r.copy()
    k = 0
    nmv = 1
    while float(np.vdot(r, r)) > atol2 and k < maxiter and k >= 0:
        rho_ = float(np.vdot(rhat, r))
        beta = rho_ / rho * alpha / omega
        p = r + beta * (p - omega * q)
        q = A(p)
        alpha = rho_ / float(np.vdot(rhat, q))
        s = r - alpha * q
        exit_early = float(np.vdot(s, s)) < atol2
        t = A(s)
        nmv += 2
        with np.errstate(all="ignore"):
            omega = float(np.vdot(t, s)) / float(np.vdot(t, t))
        if exit_early:
            x = x + alpha * p
            r = s
        else:
            x = x + (alpha * p + omega * s)
            r = s - omega * t
        k_next = -11 if (omega == 0 or alpha == 0) else k + 1
        if rho_ == 0:
            k_next = -10
        rho = rho_
        if k_next < 0:
            k = k_next
            break
        k = k_next
    if info is not None:
        info["iters"] = k
        info["matvecs"] = nmv
    return x


def gmres_restarted(A, b, tol=1e-5, atol=0.0, restart=30, maxiter=None, info=None):
    """Restarted GMRES (classical Gram-Schmidt applied twice, Givens rotations),
    x0 = 0, stop when ||r|| <= max(tol ||b||, atol).  North-star mode only: the
    reference itself never runs GMRES, so this is self-pinned."""
    b = np.asarray(b, dtype=np.float64)
    n = b.size
    if maxiter is None:
        maxiter = 10 * n
    shape = b.shape
    b = b.reshape(-1)
    x = np.zeros(n)
    target = max(tol * np.linalg.norm(b), atol)
    its = 0
    r = b.copy()
    beta = np.linalg.norm(r)
    while beta > target and its < maxiter:
        m = restart
        V = np.zeros((m + 1, n))
        H = np.zeros((m + 1, m))
        cs = np.zeros(m)
        sn = np.zeros(m)
        g = np.zeros(m + 1)
        g[0] = beta
        V[0] = r / beta
        j_used = 0
        for j in range(m):
            wv = A(V[j].reshape(shape)).reshape(-1)
            its += 1
            h = V[:j + 1] @ wv
            wv = wv - V[:j + 1].T @ h
            h2 = V[:j + 1] @ wv
            wv = wv - V[:j + 1].T @ h2
            H[:j + 1, j] = h + h2
            H[j + 1, j] = np.linalg.norm(wv)
            if H[j + 1, j] != 0:
                V[j + 1] = wv / H[j + 1, j]
            for i in range(j):
                t = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                H[i, j] = t
            den = np.hypot(H[j, j], H[j + 1, j])
            cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
            H[j, j] = den
            H[j + 1, j] = 0.0
            g[j + 1] = -sn[j] * g[j]
            g[j] = cs[j] * g[j]
            j_used = j + 1
            if abs(g[j + 1]) <= target or its >= maxiter:
                break
        y = np.linalg.solve(np.triu(H[:j_used, :j_used]), g[:j_used])
        x = x + V[:j_used].T @ y
        r = b - A(x.reshape(shape)).reshape(-1)
        beta = np.linalg.norm(r)
    if info is not None:
        info["iters"] = its
    return x.reshape(shape)


def newton_solver(f, x_init, tol=default_tolerance, max_iter=default_max_iter,
                  bicgstab_atol=1e-4, verbose=True, print_skip=1, jvp=None,
                  krylov="bicgstab", restart=30, history=None, inner=None):
    """q(x) = x - J_g(x)^{-1} g(x), g = f - id, fed to successive_approx
    (solvers.py:83-95).  ``jvp(x, v)`` must return J_f(x) v."""
    if jvp is None:
        jvp = f.jvp

    def q(x):
        gx = f(x) - x
        Jg = lambda v: jvp(x, v) - v
        info = {}
        if krylov == "bicgstab":
            b = bicgstab_jax(Jg, gx, atol=bicgstab_atol, info=info)
        else:
            b = gmres_restarted(Jg, gx, atol=bicgstab_atol, restart=restart,
                                info=info)
        if inner is not None:
            inner.append(info["iters"])
        return x - b
    return successive_approx(q, x_init, tol, max_iter, verbose, print_skip,
                             history=history)


def anderson_solver(f, x_init, tol=default_tolerance, max_iter=10000, verbose=True, history_size=10,
                    mixing_frequency=4, beta=8.0, ridge=1e-6):
    """Anderson acceleration as configured at solvers.py:98-124 (jaxopt.AndersonAcceleration with
    history_size=10, mixing_frequency=4, beta=8.0, ridge=1e-6).  PARITY UNPINNED: jaxopt is an
    un-vendored, un-pinned dependency that cannot be installed here; this restates its published
    update rule -- history of the last m iterates x_i and residuals r_i = f(x_i) - x_i, Gram matrix
    G_ij = <r_i, r_j>, alpha from [[0, 1^T], [1, G + ridge I]] [nu; alpha] = e_0, extrapolation
    sum_i alpha_i (x_i + beta r_i) once the history is full and on every mixing_frequency-th
    iteration, plain x <- f(x) otherwise; error = ||r||_2; stop when error <= tol."""
    shape = np.shape(x_init)
    x = np.asarray(x_init, dtype=np.float64).reshape(-1).copy()
    m, n = history_size, x.size
    X = np.tile(x, (m, 1))
    R = np.zeros((m, n))
    G = np.zeros((m, m))
    k, err = 0, np.inf
    while err > tol and k < max_iter:
        pos = k % m
        fx = np.asarray(f(x.reshape(shape))).reshape(-1)
        r = fx - x
        X[pos], R[pos] = x, r
        row = R @ r
        G[pos, :] = row
        G[:, pos] = row
        if k >= m and k % mixing_frequency == 0:
            H = np.zeros((m + 1, m + 1))
            H[0, 1:] = 1.0
            H[1:, 0] = 1.0
            H[1:, 1:] = G + ridge * np.eye(m)
            e0 = np.zeros(m + 1)
            e0[0] = 1.0
            alpha = np.linalg.solve(H, e0)[1:]
            x_new = alpha @ X + beta * (alpha @ R)
        else:
            x_new = fx
        err = float(np.linalg.norm(r))
        x = x_new
        k += 1
    if k == max_iter:
        print(f"Warning: Hit maximum iteration number {max_iter}")
    elif verbose:
        print(f"Iteration converged after {k} iterations")
    return x.reshape(shape), k


solvers = dict(newton=newton_solver, successive_approx=successive_approx, anderson=anderson_solver)


def solver(f, x_init, algorithm="newton", verbose=True):
    """Front end, solvers.py:154-177: defaults only, returns x* only."""
    try:
        fn = solvers[algorithm]
    except KeyError:
        print(f"Algorithm {algorithm} not found.  \n"
              "Falling back to successive approximation.\n")
        fn = successive_approx
    x_star, _ = fn(f, x_init)
    return x_star
