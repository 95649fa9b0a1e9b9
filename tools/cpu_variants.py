"""SURVEY 8(d) CPU lines for SSY (10,)^4, N = 10 000 (P = 0.8 GB fits host RAM), timed on the host cores:
(i) the GPU dense kernel's algorithm in NumPy (multi-threaded BLAS dgemv + prologue/epilogue),
(ii) the sum-factorised einsum form, and the GPU factor-form / dense kernels beside them when a GPU is present."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O

shapes = (10,) * 4
m = O.SSY()
arrays = O.discretize_ssy(m, shapes)
kop = O.KronSSY(shapes, m.params, arrays)
P, ar, ac, β, θ = O.dense_ssy(shapes, m.params, arrays)
w = np.full(int(np.prod(shapes)), 800.0)


def med(f, reps=20, warm=3):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


out = {"shapes": shapes, "N": int(w.size), "host_cpus": os.cpu_count()}
try:
    from threadpoolctl import threadpool_info
    out["blas_threads"] = [d.get("num_threads") for d in threadpool_info() if d.get("user_api") == "blas"]
except Exception:
    pass
t = med(lambda: O.dense_T(w, P, ar, ac, β, θ))
out["cpu_dense_dgemv"] = {"ms": t * 1e3, "evals_per_s": 1 / t, "GBps": 8 * w.size ** 2 / t / 1e9}
wg = w.reshape(shapes)
t = med(lambda: kop.T(wg))
out["cpu_factored_einsum"] = {"ms": t * 1e3, "evals_per_s": 1 / t}
try:
    import sdfs_via_autodiff_b200 as S
    ctx = S.Context.default()
    for storage in ("dense", "kron"):
        op = S.make_T_ssy(S.SSY(), shapes, storage=storage)
        x = ctx.full(shapes, 800.0)
        for _ in range(5):
            x = op(x)
        ctx.sync(); ctx.timer_start()
        for _ in range(200):
            x = op(x)
        ms = ctx.timer_stop_ms() / 200
        out[f"gpu_{storage}"] = {"ms": ms, "evals_per_s": 1e3 / ms}
except Exception as e:      # no GPU here: CPU lines only
    out["gpu"] = f"unavailable: {e}"
print(json.dumps(out))
