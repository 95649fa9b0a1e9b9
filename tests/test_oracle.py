"""CPU tests that pin the oracle (no GPU).

Golden data come from the reference source itself (tests/golden/make_golden.py):
the reference's loop operators T_ssy_loops / T_gcy_loops, its model defaults and
its one recorded Newton trace.
"""
import json
import os

import numpy as np
import pytest

import oracle as O


def _load(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, f"{tag}.npz"))
    n = len([k for k in z.files if k.startswith("arr")])
    arrays = tuple(z[f"arr{i}"] for i in range(n))
    return z, tuple(int(s) for s in z["shapes"]), tuple(z["params"]), arrays


@pytest.fixture(scope="module")
def facts(golden_dir):
    return json.load(open(os.path.join(golden_dir, "reference_facts.json")))


def test_model_defaults_match_reference(facts):
    s, g = O.SSY(), O.GCY()
    assert list(s.params) == facts["ssy_params"]
    assert s.θ == facts["ssy_theta"]
    assert list(g.params) == facts["gcy_params"]
    alt = O.SSY(**facts["ssy_alt_kwargs"])
    assert list(alt.params) == facts["ssy_alt_params"]
    assert alt.θ == facts["ssy_alt_theta"]


def test_rouwenhorst_properties():
    for n, rho, sig, mu in ((2, 0.9, 0.1, 0.0), (5, 0.987, 0.003, 0.0),
                            (18, 0.959, 0.0004, 0.0), (7, 0.983, 0.002, -1e-5)):
        x, P = O.rouwenhorst(n, rho, sig, mu)
        assert P.shape == (n, n) and np.all(P >= 0)
        np.testing.assert_allclose(P.sum(1), 1.0, rtol=0, atol=1e-14)
        # stationary moments of the chain equal those of the AR(1)
        vals, vecs = np.linalg.eig(P.T)
        pi = np.real(vecs[:, np.argmax(np.real(vals))])
        pi /= pi.sum()
        mean = pi @ x
        var = pi @ (x - mean) ** 2
        np.testing.assert_allclose(mean, mu / (1 - rho), rtol=1e-9, atol=1e-15)
        np.testing.assert_allclose(var, sig ** 2 / (1 - rho ** 2), rtol=1e-9)
        # conditional mean is exactly AR(1): E[x'|x] = mu + rho x
        np.testing.assert_allclose(P @ x, mu + rho * x, rtol=1e-10, atol=1e-16)


@pytest.mark.parametrize("tag", ["ssy_2345", "ssy_4765"])
def test_ssy_operator_matches_reference_loops(golden_dir, tag):
    z, shapes, params, arrays = _load(golden_dir, tag)
    # the oracle's discretiser still produces the stored factor arrays
    again = O.discretize_ssy(O.SSY(), shapes)
    for a, b in zip(arrays, again):
        np.testing.assert_array_equal(a, b)
    for w, ref in ((z["w"], z["Tw_ref"]),
                   (np.full(shapes, 800.0), z["Tw800_ref"])):
        got = O.T_ssy(w, shapes, params, arrays)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)
        P, ar, ac, β, θ = O.dense_ssy(shapes, params, arrays)
        got = O.dense_T(w.reshape(-1), P, ar, ac, β, θ).reshape(shapes)
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)
    if tag == "ssy_2345":
        got = O.T_ssy_loops(z["w"], shapes, params, arrays)
        np.testing.assert_allclose(got, z["Tw_ref"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("tag", ["gcy_232323", "gcy_234567"])
def test_gcy_operator_matches_reference_loops(golden_dir, tag):
    z, shapes, params, arrays = _load(golden_dir, tag)
    again = O.discretize_gcy(O.GCY(), shapes)
    for a, b in zip(arrays, again):
        np.testing.assert_array_equal(a, b)
    got = O.T_gcy(z["w"], shapes, params, arrays)
    np.testing.assert_allclose(got, z["Tw_ref"], rtol=1e-13, atol=0)
    if tag == "gcy_232323":
        got = O.T_gcy_loops(z["w"], shapes, params, arrays)
        np.testing.assert_allclose(got, z["Tw_ref"], rtol=1e-14, atol=0)
        P, ar, ac, β, θ = O.dense_gcy(shapes, params, arrays)
        np.testing.assert_allclose(P.sum(1), 1.0, atol=1e-13)
        for w, ref in ((z["w"], z["Tw_ref"]),
                       (np.full(shapes, 800.0), z["Tw800_ref"])):
            got = O.dense_T(w.reshape(-1), P, ar, ac, β, θ).reshape(shapes)
            np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)


def test_analytic_jvp_matches_finite_difference():
    ssy = O.SSY()
    shapes = (3, 3, 4, 4)
    arrays = O.discretize_ssy(ssy, shapes)
    op = O.KronSSY(shapes, ssy.params, arrays)
    rng = np.random.default_rng(7)
    w = 700 + 200 * rng.random(shapes)
    v = rng.standard_normal(shapes)
    h = 1e-4
    fd = (op.T(w + h * v) - op.T(w - h * v)) / (2 * h)
    np.testing.assert_allclose(op.jvp(w, v), fd, rtol=1e-7, atol=1e-9)
    P, ar, ac, β, θ = O.dense_ssy(shapes, ssy.params, arrays)
    dj = O.dense_jvp(w.reshape(-1), v.reshape(-1), P, ar, ac, β, θ)
    np.testing.assert_allclose(dj.reshape(shapes), op.jvp(w, v), rtol=1e-12)


def test_sandpit_newton_trace(facts):
    """The only numeric output recorded in the reference: Newton on SSY
    (10,10,10,10) from w0 = 800.  The first two steps are exact Newton steps up
    to the inexact BiCGSTAB solve (rel. gap < 1e-5); later steps inherit that
    inexactness (3 digits)."""
    ssy = O.SSY()
    shapes = tuple(facts["sandpit_shapes"])
    op = O.KronSSY(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
    hist, inner = [], []
    w, k = O.newton_solver(op.T, np.full(shapes, 800.0), jvp=op.jvp,
                           verbose=False, history=hist, inner=inner)
    ref = facts["sandpit_newton_errors"]
    np.testing.assert_allclose(hist[0], ref[0], rtol=1e-5)
    np.testing.assert_allclose(hist[1], ref[1], rtol=1e-5)
    np.testing.assert_allclose(hist[2], ref[2], rtol=1e-3)
    np.testing.assert_allclose(hist[3], ref[3], rtol=5e-3)
    # reference stopping quirk: last BiCGSTAB returns x0 = 0 after 0 iterations
    assert hist[-1] == 0.0 and inner[-1] == 0
    assert 5 <= k <= 8
    assert np.linalg.norm(op.T(w) - w) <= 1e-4


def test_successive_approx_counts_ssy_default_grid():
    """BASELINE.md anchor: SSY (2,3,4,5) from 800 takes 10 428 / 12 289 steps."""
    ssy = O.SSY()
    shapes = (2, 3, 4, 5)
    op = O.KronSSY(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
    w7, k7 = O.successive_approx(op.T, np.full(shapes, 800.0), verbose=False)
    w8, k8 = O.successive_approx(op.T, np.full(shapes, 800.0), tol=1e-8,
                                 verbose=False)
    assert (k7, k8) == (10428, 12289)
    np.testing.assert_allclose([w8.min(), w8.max()],
                               [741.6103853, 973.5905061], rtol=1e-9)


def test_successive_approx_semantics(capsys):
    f = lambda x: 0.5 * x + 1.0
    x, k = O.successive_approx(f, np.array([0.0]), tol=1e-3, verbose=True,
                               print_skip=2)
    out = capsys.readouterr().out
    assert "Beginning iteration" in out and "iter = 0, error = 1.0" in out
    assert f"Iteration converged after {k} iterations" in out
    assert abs(x[0] - 2.0) < 2e-3
    # iterate is accepted even on the terminating step; NaN ends the loop
    x, k = O.successive_approx(lambda x: x * np.nan, np.array([1.0]), verbose=False)
    assert k == 1 and np.isnan(x[0])
    x, k = O.successive_approx(f, np.array([0.0]), tol=0.0, max_iter=5, verbose=False)
    assert k == 5
    assert "Warning: Hit maximum iteration number 5" in capsys.readouterr().out


def test_bicgstab_and_gmres_solve_linear_system():
    rng = np.random.default_rng(3)
    n = 60
    A = np.eye(n) * 3 + 0.3 * rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    info = {}
    x = O.bicgstab_jax(lambda v: A @ v, b, tol=1e-12, info=info)
    np.testing.assert_allclose(A @ x, b, rtol=0, atol=1e-10)
    assert info["iters"] > 0
    x = O.gmres_restarted(lambda v: A @ v, b, tol=1e-12, restart=20)
    np.testing.assert_allclose(A @ x, b, rtol=0, atol=1e-9)
    # zero-iteration exit returns x0 = 0 (the Newton stopping quirk)
    x = O.bicgstab_jax(lambda v: A @ v, 1e-6 * b / np.linalg.norm(b), atol=1e-4,
                       info=info)
    assert info["iters"] == 0 and not x.any()


def test_gcy_newton_and_sdf_euler_identity():
    from oracle.sdf import e_sdf_gcy
    gcy = O.GCY()
    shapes = (3,) * 6
    arrays = O.discretize_gcy(gcy, shapes)
    op = O.KronGCY(shapes, gcy.params, arrays)
    w, k = O.newton_solver(op.T, np.full(shapes, 800.0), jvp=op.jvp,
                           bicgstab_atol=1e-11, verbose=False)
    w, _ = O.successive_approx(op.T, w, tol=1e-11, verbose=False)
    assert 456.0 < w.min() < 456.5 and 561.5 < w.max() < 562.0
    P, ar, ac, β, θ = O.dense_gcy(shapes, gcy.params, arrays)
    q_f, euler = O.sdf_dense(w, P, ar, ac, e_sdf_gcy(shapes, gcy.params, arrays),
                             β, θ)
    assert np.max(np.abs(euler)) < 1e-9
    assert np.all(q_f > 0.9) and np.all(q_f < 1.01)
    # explicit rows of Mbar integrate to q_f
    rows = np.array([0, 17, 728])
    M = O.sdf_rows(w, P, ac, e_sdf_gcy(shapes, gcy.params, arrays), β, θ, rows)
    np.testing.assert_allclose((P[rows] * M).sum(1), q_f[rows], rtol=1e-12)


def _sdf_case(model):
    from oracle import sdf as SD
    if model == "ssy":
        m, shapes = O.SSY(), (2, 3, 4, 5)
        arrays = O.discretize_ssy(m, shapes)
        kop = O.KronSSY(shapes, m.params, arrays)
        dense = O.dense_ssy(shapes, m.params, arrays)
        es, fields = SD.e_sdf_ssy(shapes, m.params, arrays), SD.state_fields_ssy(shapes, m.params, arrays)
    else:
        m, shapes = O.GCY(), (2, 3, 2, 3, 2, 3)
        arrays = O.discretize_gcy(m, shapes)
        kop = O.KronGCY(shapes, m.params, arrays)
        dense = O.dense_gcy(shapes, m.params, arrays)
        es, fields = SD.e_sdf_gcy(shapes, m.params, arrays), SD.state_fields_gcy(shapes, m.params, arrays)
    w, _ = O.newton_solver(kop.T, np.full(shapes, 800.0), jvp=kop.jvp, bicgstab_atol=1e-11, verbose=False)
    w, _ = O.successive_approx(kop.T, w, tol=1e-12, verbose=False)
    return kop, dense, es, fields, w


@pytest.mark.parametrize("model", ["ssy", "gcy"])
def test_sdf_closed_form_equals_quadrature_of_the_unintegrated_sdf(model):
    """Non-circular pin of the SDF (paper/autosdfs.tex:374-384): the closed forms q_f / e_sdf
    against a Gauss-Hermite integration of the un-integrated log M', and the pricing identity
    E[M' R_w'] = 1 evaluated from the un-integrated M' at the fixed point of the
    reference-pinned T.  A wrong sign or a missing 1/2 in e_sdf must be caught."""
    from oracle import sdf as SD
    kop, (P, ar, ac, β, θ), es, fields, w = _sdf_case(model)
    q_f, _ = O.sdf_dense(w, P, ar, ac, es, β, θ)
    q_quad, euler_quad = SD.sdf_quadrature(w, P, *fields, β, kop.γ, θ, kop.μ_c)
    np.testing.assert_allclose(q_f, q_quad, rtol=1e-12)
    assert np.max(np.abs(euler_quad)) < 1e-11
    # rule is converged: a different node count gives the same integral
    q_quad2, _ = SD.sdf_quadrature(w, P, *fields, β, kop.γ, θ, kop.μ_c, n_nodes=40)
    np.testing.assert_allclose(q_quad, q_quad2, rtol=1e-13)
    # the check has teeth: perturbed closed forms fail it
    σ_c = fields[1]
    for bad in (es * np.exp(-(kop.γ * σ_c) ** 2),            # -gamma^2 sigma_c^2 / 2 (wrong sign)
                es * np.exp(0.5 * (kop.γ * σ_c) ** 2),        # gamma^2 sigma_c^2 (missing 1/2)
                es * np.exp(2 * kop.γ * (kop.μ_c + fields[0]))):   # +gamma g_c (wrong sign)
        q_bad, _ = O.sdf_dense(w, P, ar, ac, bad, β, θ)
        assert np.max(np.abs(q_bad / q_quad - 1)) > 1e-3
    # away from the fixed point the pricing identity fails (it is not an algebraic identity)
    _, euler_off = SD.sdf_quadrature(1.01 * w, P, *fields, β, kop.γ, θ, kop.μ_c)
    assert np.max(np.abs(euler_off)) > 1e-5


def test_loglinear_matches_reference_golden(golden_dir):
    """wc_loglinear_factory of the reference (ssy_model.py:86-156, gcy_model.py:80-159), evaluated
    by the importable reference files themselves (tests/golden/make_golden.py)."""
    from oracle.loglinear import loglinear_ssy, loglinear_gcy
    g = json.load(open(os.path.join(golden_dir, "loglinear.json")))
    for tag, rec in g.items():
        if tag.startswith("ssy"):
            f = loglinear_ssy(O.SSY(**rec["kwargs"]))
        else:
            f = loglinear_gcy(O.GCY(**rec["kwargs"]))
        got = [f(tuple(x)) for x in rec["points"]]
        np.testing.assert_allclose(got, rec["values"], rtol=1e-13)


def test_anderson_oracle_converges_to_the_fixed_point():
    """Anderson with the reference's parameters (parity with jaxopt itself is unpinned)."""
    ssy = O.SSY()
    shapes = (2, 3, 4, 5)
    op = O.KronSSY(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
    w, k = O.anderson_solver(op.T, np.full(shapes, 800.0), verbose=False)
    w_sa, k_sa = O.successive_approx(op.T, np.full(shapes, 800.0), tol=1e-9, verbose=False)
    assert k < 10428                                 # fewer operator applications than plain iteration
    assert np.linalg.norm(op.T(w) - w) <= 1.5e-7
    np.testing.assert_allclose(w, w_sa, rtol=1e-7)


def test_continuous_oracle_consistency():
    """The continuous-state restatement (parity with the JAX original unpinned): quadrature rule
    moments, interpolation exactness for multilinear functions, analytic JVP vs finite differences,
    and agreement of the quadrature and a large Monte-Carlo rule."""
    from oracle.continuous import ContSSY, qnwnorm, lin_interp, build_grid_ssy
    nodes, weights = qnwnorm([5] * 4)
    np.testing.assert_allclose(weights.sum(), 1.0, rtol=1e-14)
    np.testing.assert_allclose(weights @ nodes, 0.0, atol=1e-14)
    np.testing.assert_allclose(weights @ nodes ** 2, 1.0, rtol=1e-13)
    np.testing.assert_allclose(weights @ nodes ** 4, 3.0, rtol=1e-13)
    sizes = (3, 4, 5, 6)
    grids = build_grid_ssy(O.SSY(), sizes)
    mesh = np.meshgrid(*grids, indexing="ij")
    f = 2.0 + 3e3 * mesh[0] - 1.5 * mesh[1] + 0.5 * mesh[2] * mesh[1] + 40 * mesh[3]      # multilinear
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(g[0], g[-1], 50) for g in grids])
    exact = 2.0 + 3e3 * pts[0] - 1.5 * pts[1] + 0.5 * pts[2] * pts[1] + 40 * pts[3]
    np.testing.assert_allclose(lin_interp(pts, f, grids), exact, rtol=1e-12)
    far = np.stack([np.full(3, g[-1] + 10 * (g[-1] - g[0])) for g in grids])              # nearest-edge extension
    np.testing.assert_allclose(lin_interp(far, f, grids), f[-1, -1, -1, -1], rtol=1e-14)
    n3, w3 = qnwnorm([3] * 4)
    op = ContSSY(O.SSY(), sizes, n3.T, w3)
    w = 700 + 200 * rng.random(sizes)
    v = rng.standard_normal(sizes)
    h = 1e-3
    fd = (op.T(w + h * v) - op.T(w - h * v)) / (2 * h)
    np.testing.assert_allclose(op.jvp(w, v), fd, rtol=1e-7, atol=1e-9)
    # on a smooth w the Monte-Carlo rule approaches the quadrature rule
    n5, w5 = qnwnorm([5] * 4)
    ws = 800.0 + 2e3 * mesh[0] + 20.0 * mesh[1] - 10.0 * mesh[2] + 3e3 * mesh[3]
    draws = rng.standard_normal((4, 20000))
    mc = ContSSY(O.SSY(), sizes, draws, np.full(20000, 1 / 20000))
    np.testing.assert_allclose(mc.T(ws), ContSSY(O.SSY(), sizes, n5.T, w5).T(ws), rtol=1e-2)


def test_interpolation_oracle_matches_reference_run_vectors(golden_dir):
    """utils.py:6-23 (vals_to_coords + map_coordinates(order=1, mode='nearest')) executed from the reference
    source (tests/golden/make_golden_interp.py): the oracle's multilinear interpolation reproduces it, inside the
    grid, on nodes and outside (nearest-edge extension), in 4 and 6 dimensions."""
    from oracle.continuous import lin_interp
    z = np.load(os.path.join(golden_dir, "lin_interp.npz"))
    for tag in ("d4", "d6"):
        grids = [z[f"{tag}_grid{i}"] for i in range(len(z[f"{tag}_sizes"]))]
        got = lin_interp(z[f"{tag}_x"], z[f"{tag}_vals"], grids)
        np.testing.assert_allclose(got, z[f"{tag}_y_ref"], rtol=1e-13)
