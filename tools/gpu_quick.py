"""Quick timing of the dense / factor-form operator on a few grid sizes (GPU box)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S

ctx = S.Context.default()
print(ctx.device_info())
out = {}
for shapes in [(10,) * 4, (14,) * 4, (18,) * 4]:
    N = int(np.prod(shapes))
    t0 = time.time()
    op = S.make_T_ssy(S.SSY(), shapes, storage="dense")
    ctx.sync()
    t_build = time.time() - t0
    w = ctx.full(shapes, 800.0)
    for _ in range(3):
        y = op(w)
    ctx.sync()
    reps = 20
    ctx.timer_start()
    for _ in range(reps):
        y = op(w)
    ms = ctx.timer_stop_ms() / reps
    gb = (8 * N * N + 32 * N) / 1e9
    print(f"dense {shapes} N={N} build {t_build:.2f}s  T {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s")
    out[str(shapes)] = dict(ms=ms, gbs=gb / ms * 1e3)
    if N <= 40000 or shapes == (18,) * 4:
        t0 = time.time()
        wn, k, info = S.newton_solver(op, w, verbose=False, return_info=True)
        print(f"   newton: {k} outer, inner {info['inner_iters']}, matvecs {info['matvecs']}, {time.time()-t0:.2f}s, errs {info['errors']}")
    if N <= 40000:
        t0 = time.time()
        ws, k = S.successive_approx(op, w, verbose=False)
        print(f"   SA: {k} its {time.time()-t0:.2f}s")
    del op
for shapes in [(18,) * 4, (32,) * 4, (56,) * 4]:
    N = int(np.prod(shapes))
    op = S.make_T_ssy(S.SSY(), shapes, storage="kron")
    w = ctx.full(shapes, 800.0)
    for _ in range(3):
        y = op(w)
    ctx.sync()
    reps = 20
    ctx.timer_start()
    for _ in range(reps):
        y = op(w)
    ms = ctx.timer_stop_ms() / reps
    print(f"kron {shapes} N={N}  T {ms:.3f} ms  {80*N/1e9/ms*1e3:.0f} GB/s (80N bytes)")
    t0 = time.time()
    wn, k, info = S.newton_solver(op, w, verbose=False, return_info=True)
    print(f"   newton: {k} outer, inner {info['inner_iters']}, {time.time()-t0:.2f}s")
    del op
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/quick.json", "w"))
