import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
shapes = (10,) * 4
op = S.make_sweep_operator(S.SSY(), shapes, form="dense")
g = np.linspace(5, 12, 16); p = np.linspace(1.3, 2.0, 16); b = np.linspace(0.997, 0.999, 16)
lattice = np.array([[gi, pi, bi] for gi in g for pi in p for bi in b])
for B in (512, 4096):
    t0 = time.perf_counter()
    W, it, er, info = S.sweep_solve(op, lattice[:B], algorithm="newton", return_info=True)
    ctx.sync(); dt = time.perf_counter() - t0
    Wn = np.asarray(W)
    print("newton sweep B=%d: %.2fs gemms %d outer %d..%d inner_total %d..%d nan %s w range %.2f..%.2f" % (
        B, dt, info["gemms"], it.min(), it.max(), info["inner_total"].min(), info["inner_total"].max(),
        np.isnan(Wn).any(), Wn.min(), Wn.max()), flush=True)
# corner check against single-column Newton solves
idx = [0, 15, 255, 4095]
for j in idx:
    m = S.SSY(γ=lattice[j, 0], ψ=lattice[j, 1], β=lattice[j, 2])
    op1 = S.make_T_ssy(m, shapes, storage="dense")
    w1, k1 = S.newton_solver(op1, np.full(shapes, 800.0), verbose=False)
    print("col", j, "outer sweep/single", int(it[j]), k1, "max rel diff %.2e" % float(np.max(np.abs(Wn[j] - np.asarray(w1)) / np.asarray(w1))))
