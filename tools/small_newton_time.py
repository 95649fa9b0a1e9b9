"""Newton on the reference's own small grids: wall time of the second call (first call = warm-up)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
for model, shapes, storage in (("ssy", (2, 3, 4, 5), "dense"), ("ssy", (2, 3, 4, 5), "kron"), ("gcy", (3,) * 6, "kron"), ("gcy", (3,) * 6, "dense"),
                               ("ssy", (6,) * 4, "kron"), ("ssy", (10,) * 4, "kron")):
    op = (S.make_T_ssy(S.SSY(), shapes, storage=storage) if model == "ssy" else S.make_T_gcy(S.GCY(), shapes, storage=storage))
    w0 = ctx.full(shapes, 800.0)
    S.newton_solver(op, w0, verbose=False)
    ctx.sync()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        w, k, info = S.newton_solver(op, w0, verbose=False, return_info=True)
        ctx.sync()
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ws, ks = S.successive_approx(op, w0, verbose=False)
    ctx.sync()
    tsa = time.perf_counter() - t0
    print(json.dumps({"grid": f"{model}{shapes}", "storage": storage, "N": op.N, "newton_ms": min(ts) * 1e3, "outer": int(k),
                      "applications": int(info["matvecs"]), "us_per_application": min(ts) / info["matvecs"] * 1e6,
                      "sa_ms": tsa * 1e3, "sa_iters": int(ks)}), flush=True)
print("--- storage='auto'")
for model, shapes in (("ssy", (2, 3, 4, 5)), ("gcy", (3,) * 6), ("ssy", (6,) * 4), ("ssy", (8,) * 4), ("ssy", (10,) * 4)):
    op = (S.make_T_ssy(S.SSY(), shapes) if model == "ssy" else S.make_T_gcy(S.GCY(), shapes))
    w0 = ctx.full(shapes, 800.0)
    S.newton_solver(op, w0, verbose=False); S.successive_approx(op, w0, max_iter=10, verbose=False); ctx.sync()
    t0 = time.perf_counter(); w, k, info = S.newton_solver(op, w0, verbose=False, return_info=True); ctx.sync(); tn = time.perf_counter() - t0
    t0 = time.perf_counter(); ws, ks = S.successive_approx(op, w0, verbose=False); ctx.sync(); tsa = time.perf_counter() - t0
    print(json.dumps({"grid": f"{model}{shapes}", "storage": {0: "dense", 1: "kron"}[op.storage], "N": op.N, "newton_ms": tn * 1e3,
                      "sa_ms": tsa * 1e3, "sa_us_per_iter": tsa / ks * 1e6}), flush=True)
