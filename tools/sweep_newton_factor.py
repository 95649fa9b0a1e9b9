import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
form = sys.argv[2] if len(sys.argv) > 2 else "factor"
shapes = (10,) * 4
op = S.make_sweep_operator(S.SSY(), shapes, form=form)
g = np.linspace(5, 12, 16); p = np.linspace(1.3, 2.0, 16); b = np.linspace(0.997, 0.999, 16)
lattice = np.array([[gi, pi, bi] for gi in g for pi in p for bi in b])
reps = 1 if "--once" in sys.argv else 4
for rep in range(reps):
    t0 = time.perf_counter()
    W, it, er, info = S.sweep_solve(op, lattice[:B], algorithm="newton", return_info=True)
    ctx.sync()
    print("newton sweep", form, B, "%.3f s" % (time.perf_counter() - t0), info["gemms"], "applications", flush=True)
