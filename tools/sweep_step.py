"""One-shot driver for profiling the sweep GEMM (BASELINE config 5): SSY (10,)^4, B columns."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = S.Context.default()
shapes = (10,) * 4
op = S.make_sweep_operator(S.SSY(), shapes, form="dense")
g = np.linspace(5, 12, 16); p = np.linspace(1.3, 2.0, 16); b = np.linspace(0.997, 0.999, 16)
lattice = np.array([[gi, pi, bi] for gi in g for pi in p for bi in b])[:B]
W = ctx.full((B,) + shapes, 800.0)
ctx.prof_enable(8)
for _ in range(4):
    Wn = S.sweep_apply_T(op, lattice, W)
ms, n = ctx.prof_read()
print(f"B={B} gemm kernel {ms/n:.3f} ms  {2.0*op.N*op.N*B/(ms/n)/1e9:.2f} TFLOP/s")
