"""Stand-alone dense pass back to back for several seconds vs the same pass inside the SA loop kernel
(is the loop's per-iteration time the pass at sustained clocks, or loop overhead?)."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
shapes = (18,) * 4
op = S.make_T_ssy(S.SSY(), shapes, storage="dense")
def smi():
    return subprocess.run("nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits", shell=True, capture_output=True, text=True).stdout.strip()
w = ctx.full(shapes, 800.0)
op(w); ctx.sync()
for reps in (20, 200, 600):
    ms = op.bench_pass(0, reps)
    print(f"stand-alone pass x{reps}: {ms:.3f} ms per pass [{smi()}]", flush=True)
w0 = ctx.full(shapes, 800.0)
for its in (20, 200, 600):
    ctx.sync(); t0 = time.perf_counter()
    S.successive_approx(op, w0, tol=0.0, max_iter=its, verbose=False)
    ctx.sync(); dt = time.perf_counter() - t0
    print(f"SA loop kernel x{its}: {dt / its * 1e3:.3f} ms per iteration [{smi()}]", flush=True)
ms = op.bench_pass(0, 200)
print(f"stand-alone pass x200 again: {ms:.3f} ms per pass [{smi()}]", flush=True)
