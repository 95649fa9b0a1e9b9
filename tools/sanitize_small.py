"""Small run of the round-2 kernels for compute-sanitizer (memcheck / racecheck): fused factor-form apply (T, JVP, SDF, P),
one-CTA factor-form SA, cooperative SA / Newton loops on a mid-size grid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
for shapes in ((3, 4, 5, 6), (12, 3, 10, 9), (9, 10, 11, 12)):
    op = S.make_T_ssy(S.SSY(), shapes, storage="kron")
    rng = np.random.default_rng(0)
    w = 500 + 300 * rng.random(shapes)
    v = rng.standard_normal(shapes)
    np.asarray(op(w)); np.asarray(op.jvp(w, v)); np.asarray(op.apply_P(w)); op.sdf(w)
    S.successive_approx(op, w, max_iter=20, tol=0.0, verbose=False)
    S.newton_solver(op, np.full(shapes, 800.0), max_iter=2, verbose=False)
g = S.make_T_gcy(S.GCY(), (3, 2, 3, 2, 3, 2), storage="kron")
np.asarray(g(np.full((3, 2, 3, 2, 3, 2), 700.0)))
S.successive_approx(g, np.full((3, 2, 3, 2, 3, 2), 800.0), max_iter=20, tol=0.0, verbose=False)
print("sanitize_small done")
