import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import sdfs_via_autodiff_b200 as S
shapes = (2, 3, 4, 5)
ssy = O.SSY(); kop = O.KronSSY(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
w_ref, k_ref = O.successive_approx(kop.T, np.full(shapes, 800.0), tol=1e-8, verbose=False)
op = S.make_T_ssy(S.SSY(), shapes, storage="dense")
for v in ("1", "0", "161", "320"):
    os.environ["SDFS_SMALL_VARIANT"] = v
    S.successive_approx(op, np.full(shapes, 800.0), tol=1e-8, verbose=False)
    t0 = time.perf_counter()
    w, k = S.successive_approx(op, np.full(shapes, 800.0), tol=1e-8, verbose=False)
    dt = time.perf_counter() - t0
    print(f"variant {v}: iters {k} (ref {k_ref}) {dt*1e3:.2f} ms  {dt/k*1e6:.3f} us/iter  maxrel {np.max(np.abs(np.asarray(w)-w_ref)/w_ref):.2e}", flush=True)
