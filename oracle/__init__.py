"""CPU oracle for the wealth-consumption hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm
(jstac/sdfs_via_autodiff, hot path of code/solvers.py with the SSY and GCY
discretised models).  It exists to *check* the CUDA product, never to serve it:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``sdfs_via_autodiff_b200/`` imports it, and the product raises if its CUDA
library is missing -- there is no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * operator T (SSY, GCY): PINNED against the reference's own loop oracles
    ``T_ssy_loops`` / ``T_gcy_loops`` executed from the reference source
    (tests/golden/make_golden.py -> tests/golden/*.npz).
  * model defaults / theta: PINNED against the importable reference model files.
  * Rouwenhorst discretiser, BiCGSTAB recurrence, Newton loop: pinned only
    through the single recorded Newton trace of the reference
    (code/ssy/discrete/sandpit.ipynb), reproduced to 3-7 digits; the
    third-party algorithms (quantecon.rouwenhorst, jax bicgstab) are restated
    from their published form because neither library is installable here.
  * interpolation of the continuous-state rows (utils.py:6-23): PINNED against the reference's own
    functions executed from source (tests/golden/make_golden_interp.py -> lin_interp.npz).
  * SDF: the reference has no SDF code.  The closed forms (e_sdf, q_f, M-bar) are PINNED to an
    independent Gauss-Hermite integration of the paper's un-integrated log M' (oracle/sdf.py::
    sdf_quadrature, paper/autosdfs.tex:374-384), including the pricing identity E[M' R_w'] = 1 at the
    fixed point of the reference-pinned T; a wrong sign or a missing 1/2 fails the check.
  * converged w*, iteration counts, GCY solutions, Anderson vs jaxopt, Monte-Carlo draws: parity
    unpinned (self-pinned by this oracle).
"""
from .models import SSY, GCY                                    # noqa: F401
from .rouwenhorst import rouwenhorst                            # noqa: F401
from .discretize import discretize_ssy, discretize_gcy          # noqa: F401
from .operators import (T_ssy_loops, T_gcy_loops, T_ssy, T_gcy,  # noqa: F401
                        dense_ssy, dense_gcy, dense_T, dense_jvp,
                        KronSSY, KronGCY)
from .solvers import (successive_approx, newton_solver, solver,  # noqa: F401
                      bicgstab_jax, gmres_restarted, solvers, anderson_solver)
from .sdf import sdf_dense, sdf_rows                             # noqa: F401
