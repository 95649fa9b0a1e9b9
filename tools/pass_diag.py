import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ctx = S.Context.default()
shapes = (18,) * 4
op = S.make_T_ssy(S.SSY(), shapes, storage="dense")
N = op.N
gb = (8 * N * N + 32 * N) / 1e9
w = ctx.full(shapes, 800.0)
y = op(w); ctx.sync()
def smi():
    return subprocess.run("nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits", shell=True, capture_output=True, text=True).stdout.strip()
for mode, name in ((0, "T epilogue"), (3, "plain Px"), (0, "T epilogue again")):
    for reps in (5, 40, 160):
        ms = op.bench_pass(mode, reps)
        print(f"back-to-back {name:18s} reps {reps:4d}: {ms:.3f} ms  {gb/ms*1e3:.0f} GB/s   [{smi()}]", flush=True)
# python-driven loop, as bench.py does
for reps in (20, 150):
    ctx.prof_enable(reps)
    w = ctx.full(shapes, 800.0)
    ctx.timer_start()
    for _ in range(reps):
        w = op(w)
    ms = ctx.timer_stop_ms() / reps
    kms, n = ctx.prof_read()
    print(f"python loop reps {reps}: step {ms:.3f} ms, kernel {kms/n:.3f} ms  [{smi()}]", flush=True)
# synthetic x (same magnitudes as the probe) to test data dependence
x = ctx.asarray(1.0 + (np.arange(N) % 13).astype(np.float64))
op.apply_P(x); ctx.sync()
print(f"plain Px, probe-like x: {op.bench_pass(3, 80):.3f} ms", flush=True)
