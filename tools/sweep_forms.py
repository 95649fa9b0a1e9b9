"""BASELINE config 5 on one GPU, both forms of S = P V: the tensor-core GEMM against the stored P and the
batched factor-form contraction.  Newton sweep of B parameter sets of SSY (10,)^4; optionally a larger grid
(factor form only: no P)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
g = np.linspace(5, 12, 16); p = np.linspace(1.3, 2.0, 16); b = np.linspace(0.997, 0.999, 16)
lattice = np.array([[gi, pi, bi] for gi in g for pi in p for bi in b])
out = {}
res = {}
for form in ("factor", "dense"):
    shapes = (10,) * 4
    op = S.make_sweep_operator(S.SSY(), shapes, form=form)
    for B in (512, 4096):
        W0 = ctx.full((B,) + shapes, 800.0)
        for _ in range(2):
            S.sweep_apply_T(op, lattice[:B], W0)
        ctx.sync(); ctx.prof_enable(8)
        for _ in range(5):
            S.sweep_apply_T(op, lattice[:B], W0)
        ms, n = ctx.prof_read(); ctx.prof_enable(0)
        S.sweep_solve(op, lattice[:8], algorithm="newton")
        ctx.sync(); t0 = time.perf_counter()
        W, it, er, info = S.sweep_solve(op, lattice[:B], algorithm="newton", return_info=True)
        ctx.sync(); dt = time.perf_counter() - t0
        res[(form, B)] = np.asarray(W)
        out[f"{form}_B{B}"] = dict(PV_step_ms=ms / n, newton_seconds=dt, applications=int(info["gemms"]),
                                   outer_min=int(it.min()), outer_max=int(it.max()))
        print(form, B, out[f"{form}_B{B}"], flush=True)
    del op
for B in (512, 4096):
    out[f"max_rel_diff_B{B}"] = float(np.max(np.abs(res[("factor", B)] - res[("dense", B)]) / res[("dense", B)]))
# a grid the dense form cannot hold: (20,)^4 = 160 000 states (P would be 205 GB), 1024 parameter sets
shapes = (20,) * 4
op = S.make_sweep_operator(S.SSY(), shapes, form="factor")
B = 1024
S.sweep_solve(op, lattice[:8], algorithm="newton")
ctx.sync(); t0 = time.perf_counter()
W, it, er, info = S.sweep_solve(op, lattice[:B], algorithm="newton", return_info=True)
ctx.sync(); dt = time.perf_counter() - t0
out["factor_20x4_B1024"] = dict(newton_seconds=dt, applications=int(info["gemms"]), outer_min=int(it.min()),
                                outer_max=int(it.max()), nan=bool(np.isnan(np.asarray(W)).any()))
print(json.dumps(out), flush=True)
