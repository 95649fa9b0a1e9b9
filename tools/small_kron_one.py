"""One SA solve on the reference's default GCY grid (3,)^6 in factor form (ncu target for k_sa_kron_small)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdfs_via_autodiff_b200 as S
op = S.make_T_gcy(S.GCY(), (3,) * 6, storage="kron")
w, k = S.successive_approx(op, np.full((3,) * 6, 800.0), max_iter=2000, tol=0.0, verbose=False)
print("iterations", k)
