"""Factor-form apply and Newton timings at the BASELINE configs[3] sizes (one GPU).
usage: python tools/kron_time.py [18 32 56 ...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S

ctx = S.Context.default()
flush = ctx.empty((1 << 25,))
out = {}
for n in [int(a) for a in sys.argv[1:]] or [18, 32, 56]:
    shapes = (n,) * 4
    op = S.make_T_ssy(S.SSY(), shapes, storage="kron")
    N = op.N
    w = ctx.full(shapes, 800.0)
    for _ in range(3):
        w = op(w)
    ctx.sync()
    ctx.timer_start()
    for _ in range(20):
        w = op(w)
    chain = ctx.timer_stop_ms() / 20
    tot = 0.0
    for _ in range(10):
        flush.fill(0.0)
        ctx.timer_start()
        w = op(w)
        tot += ctx.timer_stop_ms()
    S.newton_solver(op, ctx.full(shapes, 800.0), verbose=False)
    ctx.sync()
    t0 = time.perf_counter()
    wn, k, info = S.newton_solver(op, ctx.full(shapes, 800.0), verbose=False, return_info=True)
    ctx.sync()
    dt = time.perf_counter() - t0
    out[n] = dict(N=N, T_ms_chained=chain, T_ms_flushed=tot / 10, GBps_80N=80 * N / (tot / 10) / 1e6,
                  newton_s=dt, outer=int(k), apps=int(info["matvecs"]), ms_per_app=dt / info["matvecs"] * 1e3)
    print(json.dumps({n: out[n]}), flush=True)
    del op, w, wn
