"""CPU tests: the C-ABI library loads and exports every declared symbol, and the
host-side mirror of the reference interface behaves like the reference (no GPU,
no compute calls)."""
import json
import os
import re

import numpy as np
import pytest

import sdfs_via_autodiff_b200 as S
from sdfs_via_autodiff_b200 import _lib


def test_library_exports_every_declared_symbol():
    declared = _lib.declared_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} declared in include/sdfs_b200.h but not exported"
    # and every bound signature is declared in the header
    assert set(_lib._SIGS) <= set(declared)
    assert _lib.lib.sdfs_abi_version() == 1
    assert b"sm_100a" in _lib.lib.sdfs_version_string()


def test_header_cites_reference_for_each_group():
    text = open(_lib.HEADER).read()
    for cite in ("solvers.py:19-48", "solvers.py:51-95", "ssy_wc_ratio.py:23-79",
                 "gcy_wc_ratio.py:31-131", "temp_ssy.py:204-216", "autosdfs.tex:374-384"):
        assert cite in text


def test_no_cpu_fallback_without_gpu():
    import ctypes as C
    h = C.c_void_p()
    n = C.c_int()
    rc = _lib.lib.sdfs_ctx_create(10_000, C.byref(h))
    assert rc != 0 and not h.value
    assert _lib.lib.sdfs_last_error(None)
    # product package never imports the oracle
    pkg = os.path.dirname(S.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn
            assert "import torch" not in src and "import triton" not in src, fn


def test_models_match_reference_defaults(golden_dir):
    facts = json.load(open(os.path.join(golden_dir, "reference_facts.json")))
    s, g = S.SSY(), S.GCY()
    assert [float(x) for x in s.params] == facts["ssy_params"]
    assert float(s.θ) == facts["ssy_theta"]
    assert [float(x) for x in g.params] == facts["gcy_params"]
    assert not hasattr(g, "θ")                      # like the reference
    alt = S.SSY(**facts["ssy_alt_kwargs"])
    assert [float(x) for x in alt.params] == facts["ssy_alt_params"]


def test_solver_front_end_semantics(capsys):
    import importlib
    sv = importlib.import_module("sdfs_via_autodiff_b200.solvers")   # (the package attribute `solvers` is the dict)
    f = lambda x: 0.5 * x + 1.0                     # a generic callable is NOT iterated on the host
    with pytest.raises(TypeError, match="no CPU fallback"):
        S.successive_approx(f, np.array([0.0]), verbose=False)
    with pytest.raises(TypeError, match="no CPU fallback"):
        S.newton_solver(f, np.array([0.0]), verbose=False)
    assert set(S.solvers) == {"newton", "anderson", "gd", "successive_approx"}
    # a failure inside the closure that is not caused by the probe (bad shapes, out of memory, CUDA
    # errors while the operator is built) propagates instead of being reported as "not an operator"

    def broken(w):
        raise ValueError("factor array has shape (3,), expected (4,)")
    with pytest.raises(ValueError, match="factor array has shape"):
        S.successive_approx(broken, np.array([0.0]), verbose=False)
    # unknown algorithm: the reference's message, then successive approximation (solvers.py:164-172)
    with pytest.raises(TypeError):
        S.solver(f, np.array([0.0]), algorithm="no-such-algo")
    out = capsys.readouterr().out
    assert "Algorithm no-such-algo not found." in out and "Falling back to successive approximation." in out
    with pytest.raises(TypeError, match="no CPU fallback"):
        S.solvers["anderson"](f, np.array([0.0]))
    with pytest.raises(NotImplementedError):
        S.solvers["gd"](f, np.array([0.0]))
    assert S.default_tolerance == 1e-7 and S.default_max_iter == 1000000
    # the printed trace is rebuilt from the device-side history exactly as the reference prints it
    hist = [1.0, 0.25, 0.0625]                      # errors of iterations 0, 2, 4 (print_skip = 2)
    sv._report(hist, 5, 100, True, 2, stride=2)
    out = capsys.readouterr().out
    assert out == ("iter = 0, error = 1.0\niter = 2, error = 0.25\niter = 4, error = 0.0625\n"
                   "Iteration converged after 5 iterations\n")
    sv._report(None, 7, 7, False, 1000)
    assert capsys.readouterr().out == "Warning: Hit maximum iteration number 7\n"


def test_host_loglinear_factory_matches_reference_golden(golden_dir):
    """Host mirror of wc_loglinear_factory (own root finder instead of scipy.brentq): agrees with
    the values produced by the reference files to brentq's own tolerance (xtol 2e-12)."""
    from sdfs_via_autodiff_b200.ssy_model import wc_loglinear_factory as ssy_ll
    from sdfs_via_autodiff_b200.gcy_model import wc_loglinear_factory as gcy_ll
    g = json.load(open(os.path.join(golden_dir, "loglinear.json")))
    for tag, rec in g.items():
        f = ssy_ll(S.SSY(**rec["kwargs"])) if tag.startswith("ssy") else gcy_ll(S.GCY(**rec["kwargs"]))
        got = [f(tuple(x)) for x in rec["points"]]
        np.testing.assert_allclose(got, rec["values"], rtol=1e-9)
        assert set(f.coeffs) >= {"A0", "Ah_λ", "Ah_c", "Ah_z", "Az", "qbar"}


def test_sweep_front_end_argument_checks():
    """Host-side validation happens before any device call (runs without a GPU)."""
    with pytest.raises(ValueError):
        S.make_sweep_operator(S.SSY(), (2, 3, 4, 5), form="gemm")
    from sdfs_via_autodiff_b200.sweep import column_slice, _prefs
    # column sharding covers every column exactly once, ragged tails included
    for B, G in ((4096, 8), (5, 2), (3, 8), (1, 1)):
        got = [column_slice(B, G, r) for r in range(G)]
        assert got[0][0] == 0 and got[-1][1] == B
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
    with pytest.raises(ValueError):
        _prefs(np.ones((3, 2)))
    assert _prefs([[8.89, 1.97, 0.999]]).shape == (1, 3)


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) on a tiny grid: one JSON line with
    the contract's keys, every timed step a FULL evaluation of the reference's broadcast-sum (nothing extrapolated),
    checked against the sum-factorised oracle; ranks other than 0 print nothing."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--shapes", "3,3,4,5", "--steps", "2",
           "--warmup", "1", "--cpu-threads", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["dtype"] == "f64"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["full_evaluations"] is True and "nothing extrapolated" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value = full evaluations per second: ms_per_step is the time of ONE evaluation
    np.testing.assert_allclose(d["value"], 1e3 / d["ms_per_step"], rtol=1e-9)
    assert d["cpu_factored"]["value"] > 0
    # non-zero ranks of a torchrun launch stay silent and exit 0
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
