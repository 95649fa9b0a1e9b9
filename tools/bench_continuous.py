"""Continuous-state operator measurements (SSY default grid of the reference: 10,10,10,20 states,
d = 5 Gauss-Hermite nodes per dimension = 625 nodes) on one B200, CPU oracle beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from oracle.continuous import ContSSY
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
out = {}
for sizes, d in (((10, 10, 10, 20), 5), ((15, 15, 15, 15), 5)):
    grids, T = S.make_T_continuous(S.SSY(), sizes, d=d)
    N = int(np.prod(sizes)); Q = d ** 4
    w = ctx.full(sizes, 800.0)
    for _ in range(3):
        w = T(w)
    ctx.sync(); ctx.timer_start()
    reps = 20
    for _ in range(reps):
        w = T(w)
    ms = ctx.timer_stop_ms() / reps
    v = ctx.full(sizes, 1.0)
    T.jvp(w, v); ctx.sync(); ctx.timer_start()
    for _ in range(reps):
        T.jvp(w, v)
    jms = ctx.timer_stop_ms() / reps
    r = dict(N=N, Q=Q, T_ms=ms, interp_pow_per_s=N * Q / ms * 1e3, fused_T_jvp_ms=jms)
    for algo, kw in (("newton", {}), ("anderson", {}), ("successive_approx", {})):
        t0 = time.perf_counter()
        ws, k = S.solvers[algo](T, ctx.full(sizes, 800.0), verbose=False, **kw)
        ctx.sync()
        r[algo] = dict(iters=int(k), seconds=time.perf_counter() - t0)
        wsn = np.asarray(ws)
        r[algo]["w_range"] = [float(wsn.min()), float(wsn.max())]
        r[algo]["residual_max"] = float(np.max(np.abs(np.asarray(T(ws)) - wsn)))
    if sizes == (10, 10, 10, 20):
        nodes, weights = S.gauss_hermite_normal(d, 4)
        ref = ContSSY(O.SSY(), sizes, nodes, weights)
        w0 = np.full(sizes, 800.0)
        t0 = time.perf_counter(); Tw = ref.T(w0); cpu = time.perf_counter() - t0
        r["cpu_oracle_T_s"] = cpu
        r["T_maxrel_vs_oracle"] = float(np.max(np.abs(np.asarray(T(w0)) - Tw) / Tw))
    out[str(sizes)] = r
    print(json.dumps({str(sizes): r}), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/continuous.json", "w"), indent=1)
