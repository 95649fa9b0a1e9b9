"""One rank's row slab of the 88 GB dense operator on ONE GPU: pass time per row with and without the last-wave
split (SDFS_DENSE_TAIL=0/1 is read once per process, so run twice):
    python tools/dense_tail.py [ranks]      slab = rows of rank 0 of `ranks` (default 8) of the (18,)^4 operator"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
ranks = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = S.Context.default()
N = 18 ** 4
for rows in ((N + ranks - 1) // ranks, ((N // ranks) // (8 * 148)) * 8 * 148):
    P = ctx.full((rows, N), 1.0 / N)
    op = S.WCOperator.from_dense(P, np.ones(N), np.ones(N), 0.99, -5.0, row_range=(0, rows))
    w = ctx.full((N,), 800.0)
    out = np.asarray(op(w))[:rows]
    ref = 1.0 + 0.99 * (800.0 ** -5.0) ** (1.0 / -5.0)
    ok = np.allclose(out, ref, rtol=1e-12)
    ms = min(op.bench_pass(0, 50) for _ in range(3))
    groups = (rows + 7) // 8
    print(f"tail={os.environ.get('SDFS_DENSE_TAIL', '1')} rows {rows} ({groups} groups = {groups / 148:.2f} waves): {ms * 1e3:.1f} us per pass, "
          f"{ms * 1e3 / rows * 1e3:.2f} ns per row, {rows * N * 8 / ms * 1e-6:.0f} GB/s, values ok {ok}", flush=True)
    del op, P
