"""SA sweep of the (gamma, psi, beta) lattice in factor form on one GPU: 512 columns (one GPU's share of 8) and all 4096."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
shapes = (10,) * 4
op = S.make_sweep_operator(S.SSY(), shapes, form="factor")
g = np.linspace(5, 12, 16); p = np.linspace(1.3, 2.0, 16); b = np.linspace(0.997, 0.999, 16)
lattice = np.array([[gi, pi, bi] for gi in g for pi in p for bi in b])
S.sweep_solve(op, lattice[:8], max_iter=50)
for B in (512, 4096):
    ctx.sync(); t0 = time.perf_counter()
    W, it, er = S.sweep_solve(op, lattice[:B])
    ctx.sync(); dt = time.perf_counter() - t0
    it = np.asarray(it)
    print(f"SA sweep factor form B={B}: {dt:.2f} s, steps per column {it.min()}..{it.max()}, {dt / it.max() * 1e3:.3f} ms per step, "
          f"max err {float(np.max(er)):.2e}", flush=True)
