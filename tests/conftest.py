import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_library():
    """The shared library is git-ignored: in a fresh checkout build it before the package is
    imported (nvcc cross-compiles sm_100a without a GPU).  Without nvcc the import below fails
    loudly -- there is no CPU fallback to test."""
    import importlib.util
    import shutil
    spec = importlib.util.spec_from_file_location("sdfs_build", os.path.join(ROOT, "sdfs_via_autodiff_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    have_nvcc = os.path.exists("/usr/local/cuda/bin/nvcc") or shutil.which("nvcc")
    if have_nvcc and (not os.path.exists(mod.LIB) or mod.needs_build()):
        mod.build()


_ensure_library()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
