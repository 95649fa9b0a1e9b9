"""SA on mid-size grids in factor form: shared-memory-resident one-CTA loop vs the cooperative loop kernel.
usage: python tools/small_kron_time.py   (SDFS_SMALL_KRON_MAX=0 disables the one-CTA path)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import sdfs_via_autodiff_b200 as S

out = {}
cases = [("ssy", (2, 3, 4, 5)), ("ssy", (3, 4, 5, 6)), ("gcy", (3,) * 6), ("ssy", (6, 6, 6, 6)), ("ssy", (8, 8, 8, 8)),
         ("gcy", (4,) * 6), ("ssy", (9, 9, 9, 9))]
for model, shapes in cases:
    if model == "ssy":
        op = S.make_T_ssy(S.SSY(), shapes, storage="kron")
        m = O.SSY(); kop = O.KronSSY(shapes, m.params, O.discretize_ssy(m, shapes))
    else:
        op = S.make_T_gcy(S.GCY(), shapes, storage="kron")
        m = O.GCY(); kop = O.KronGCY(shapes, m.params, O.discretize_gcy(m, shapes))
    w0 = np.full(shapes, 800.0)
    S.successive_approx(op, w0, tol=1e-7, max_iter=10, verbose=False)
    t0 = time.perf_counter()
    w, k = S.successive_approx(op, w0, tol=1e-7, verbose=False)
    dt = time.perf_counter() - t0
    w_ref, k_ref = O.successive_approx(kop.T, w0, tol=1e-7, verbose=False)
    out[f"{model}{shapes}"] = dict(N=op.N, iters=int(k), iters_oracle=int(k_ref), seconds=dt, us_per_iter=dt / k * 1e6,
                                  max_rel=float(np.max(np.abs(np.asarray(w) / w_ref - 1))))
    print(json.dumps({f"{model}{shapes}": out[f"{model}{shapes}"]}), flush=True)
