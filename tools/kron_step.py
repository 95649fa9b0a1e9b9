"""Factor-form apply timing: python tools/kron_step.py <n> [ssy|gcy]  -> (n,)^4 SSY or (n,)^6 GCY."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 56
model = sys.argv[2] if len(sys.argv) > 2 else "ssy"
ctx = S.Context.default()
if model == "gcy":
    shapes = (n,) * 6
    op = S.make_T_gcy(S.GCY(), shapes, storage="kron")
else:
    shapes = (n,) * 4
    op = S.make_T_ssy(S.SSY(), shapes, storage="kron")
w = ctx.full(shapes, 800.0)
for _ in range(3):
    w = op(w)
ctx.sync(); ctx.timer_start()
for _ in range(10):
    w = op(w)
print(f"kron {model} {shapes} N={op.N}: {ctx.timer_stop_ms()/10:.3f} ms per apply")
if "--newton" in sys.argv:
    w0 = ctx.full(shapes, 800.0)
    S.newton_solver(op, w0, max_iter=1, verbose=False)
    ctx.sync(); t0 = time.perf_counter()
    ws, k, info = S.newton_solver(op, w0, verbose=False, return_info=True)
    ctx.sync()
    print(f"  newton: {time.perf_counter()-t0:.3f} s, {k} outer, {info['matvecs']} applications")
