import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 56
ctx = S.Context.default()
shapes = (n,) * 4
op = S.make_T_ssy(S.SSY(), shapes, storage="kron")
w = ctx.full(shapes, 800.0)
for _ in range(3):
    w = op(w)
ctx.sync(); ctx.timer_start()
for _ in range(10):
    w = op(w)
print(f"kron {shapes}: {ctx.timer_stop_ms()/10:.3f} ms per apply")
