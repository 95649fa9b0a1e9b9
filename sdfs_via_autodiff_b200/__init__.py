"""B200-native solver for the wealth-consumption-ratio operator, its fixed-point
loops and the SDF (hot path of jstac/sdfs_via_autodiff code/solvers.py with the
SSY and GCY discretised models).

Python host + ctypes/DLPack over libsdfs_b200.so (hand-written CUDA for sm_100a).
Importing this package loads the shared library; if it is missing the import fails
-- there is no CPU fallback and no PyTorch/Triton path.
"""
from ._lib import lib, SdfsError, LIB_PATH                      # noqa: F401
from .device import Context, DeviceArray, from_dlpack           # noqa: F401
from .ssy_model import SSY                                       # noqa: F401
from .gcy_model import GCY                                       # noqa: F401
from .operator import WCOperator, Factors                        # noqa: F401
from .ssy_wc_ratio import discretize_ssy, T_ssy, T_ssy_loops, make_T_ssy, test_compute_wc_ratio_ssy   # noqa: F401
from .gcy_wc_ratio import discretize_gcy, T_gcy, T_gcy_loops, make_T_gcy, test_compute_wc_ratio_gcy   # noqa: F401
from .solvers import (successive_approx, newton_solver, anderson_solver, solver, solvers,                 # noqa: F401
                      default_tolerance, default_max_iter)
from .sdf import solve_ssy, solve_gcy, SDFResult                 # noqa: F401
from .sweep import make_sweep_operator, sweep_apply_T, sweep_solve                      # noqa: F401
from .loglinear import loglinear_guess                                               # noqa: F401
from .continuous import (build_grid, T_fun_factory, make_T_continuous, wc_ratio_continuous,   # noqa: F401
                         gauss_hermite_normal, construct_wstar_callable, lin_interp, save_wstar, compare_T_factories)
