"""Per-config measurements for BASELINE.json configs C1..C5 on one B200 (profiles/ evidence).
CUDA-event timings where a kernel is timed, wall-clock for whole solves."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
import sdfs_via_autodiff_b200 as S

ctx = S.Context.default()
out = {}
which = set(sys.argv[1:]) or {"c1", "c2", "c3", "c4", "c5"}


def timed_T(op, reps):
    w = ctx.full(op.shapes, 800.0)
    for _ in range(3):
        w = op(w)
    ctx.sync()
    ctx.prof_enable(reps)
    ctx.timer_start()
    for _ in range(reps):
        w = op(w)
    ms = ctx.timer_stop_ms() / reps
    kms, n = ctx.prof_read()
    ctx.prof_enable(0)
    return ms, (kms / n if n else None)


if "c1" in which:
    shapes = (2, 3, 4, 5)
    ssy = O.SSY()
    kop = O.KronSSY(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
    r = {}
    for storage in ("dense", "kron"):
        op = S.make_T_ssy(S.SSY(), shapes, storage=storage)
        for tol in (1e-7, 1e-8):
            S.successive_approx(op, np.full(shapes, 800.0), tol=tol, verbose=False)
            t0 = time.perf_counter()
            w, k = S.successive_approx(op, np.full(shapes, 800.0), tol=tol, verbose=False)
            dt = time.perf_counter() - t0
            r[f"sa_{storage}_tol{tol:g}"] = dict(iters=int(k), seconds=dt, us_per_iter=dt / k * 1e6,
                                                 evals_per_s=k / dt)
        t0 = time.perf_counter()
        w, k, info = S.newton_solver(op, np.full(shapes, 800.0), verbose=False, return_info=True)
        r[f"newton_{storage}"] = dict(outer=int(k), seconds=time.perf_counter() - t0, matvecs=info["matvecs"])
    t0 = time.perf_counter()
    w_ref, k_ref = O.successive_approx(kop.T, np.full(shapes, 800.0), tol=1e-8, verbose=False)
    dt = time.perf_counter() - t0
    r["cpu_oracle_factored_sa_tol1e-08"] = dict(iters=k_ref, seconds=dt, us_per_iter=dt / k_ref * 1e6)
    out["C1 SSY (2,3,4,5) N=120"] = r
    print(json.dumps({"C1": r}), flush=True)

if "c2" in which:
    shapes = (18,) * 4
    op = S.make_T_ssy(S.SSY(), shapes, storage="dense")
    N = op.N
    ms, kms = timed_T(op, 20)
    r = dict(T_ms=ms, kernel_ms=kms, GBps=(8 * N * N + 32 * N) / kms / 1e6)
    for kry in ("bicgstab", "gmres"):
        for tol in (1e-7, 1e-8):
            t0 = time.perf_counter()
            w, k, info = S.newton_solver(op, ctx.full(shapes, 800.0), tol=tol, krylov=kry, verbose=False,
                                         return_info=True)
            ctx.sync()
            dt = time.perf_counter() - t0
            res = float(np.max(np.abs(np.asarray(op(w)) - np.asarray(w))))
            r[f"newton_{kry}_tol{tol:g}"] = dict(outer=int(k), inner=[int(x) for x in info["inner_iters"]],
                                                 applications=int(info["matvecs"]), seconds=dt,
                                                 apps_per_s=info["matvecs"] / dt, residual=res)
    # JVP: fused T+JVP single pass
    w = ctx.full(shapes, 800.0); v = ctx.full(shapes, 1.0)
    for _ in range(2):
        op.jvp(w, v)
    ctx.sync(); ctx.timer_start()
    for _ in range(10):
        op.jvp(w, v)
    jms = ctx.timer_stop_ms() / 10
    r["fused_T_JVP_ms"] = jms
    r["fused_T_JVP_GBps"] = (8 * N * N + 48 * N) / jms / 1e6
    out["C2 SSY (18,)^4 N=104976 dense 88.2GB"] = r
    print(json.dumps({"C2": r}), flush=True)
    del op

if "c3" in which:
    shapes = (7,) * 6
    t0 = time.perf_counter()
    op = S.make_T_gcy(S.GCY(), shapes, storage="dense")
    ctx.sync()
    build = time.perf_counter() - t0
    N = op.N
    ms, kms = timed_T(op, 10)
    r = dict(build_s=build, T_ms=ms, kernel_ms=kms, GBps=(8 * N * N + 32 * N) / kms / 1e6)
    t0 = time.perf_counter()
    w, k, info = S.newton_solver(op, ctx.full(shapes, 800.0), verbose=False, return_info=True)
    ctx.sync()
    dt = time.perf_counter() - t0
    r["newton_bicgstab"] = dict(outer=int(k), inner=[int(x) for x in info["inner_iters"]],
                                applications=int(info["matvecs"]), seconds=dt)
    ctx.timer_start()
    qf, eu = op.sdf(w)
    sms = ctx.timer_stop_ms()
    wn = np.asarray(w)
    r["sdf_pass_ms"] = sms
    r["sdf_GBps"] = (8 * N * N + 64 * N) / sms / 1e6
    r["w_range"] = [float(wn.min()), float(wn.max())]
    r["euler_max_abs"] = float(np.max(np.abs(np.asarray(eu))))
    r["qf_range"] = [float(np.asarray(qf).min()), float(np.asarray(qf).max())]
    out["C3 GCY (7,)^6 N=117649 dense 110.7GB"] = r
    print(json.dumps({"C3": r}), flush=True)
    del op

if "c4" in which:
    r = {}
    for shapes in ((18,) * 4, (32,) * 4, (56,) * 4):
        op = S.make_T_ssy(S.SSY(), shapes, storage="kron")
        N = op.N
        w = ctx.full(shapes, 800.0)
        for _ in range(3):
            w = op(w)
        ctx.sync(); ctx.timer_start()
        for _ in range(20):
            w = op(w)
        ms = ctx.timer_stop_ms() / 20
        t0 = time.perf_counter()
        wn, k, info = S.newton_solver(op, ctx.full(shapes, 800.0), verbose=False, return_info=True)
        ctx.sync()
        dt = time.perf_counter() - t0
        r[str(shapes)] = dict(N=N, T_ms=ms, GBps_80N=80 * N / ms / 1e6, newton_outer=int(k),
                              newton_inner=[int(x) for x in info["inner_iters"]],
                              newton_applications=int(info["matvecs"]), newton_seconds=dt)
        del op
    out["C4 factor-form SSY"] = r
    print(json.dumps({"C4": r}), flush=True)

if "c5" in which:
    shapes = (10,) * 4
    op = S.make_sweep_operator(S.SSY(), shapes, form="dense")
    N = op.N
    g = np.linspace(5, 12, 16); p = np.linspace(1.3, 2.0, 16); b = np.linspace(0.997, 0.999, 16)
    lattice = np.array([[gi, pi, bi] for gi in g for pi in p for bi in b])
    r = {}
    for B in (512, 4096):
        prefs = lattice[:B]
        W = ctx.full((B,) + shapes, 800.0)
        for _ in range(2):
            S.sweep_apply_T(op, prefs, W)
        ctx.sync()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            Wn = S.sweep_apply_T(op, prefs, W)
        ctx.sync()
        dt = (time.perf_counter() - t0) / reps
        ctx.prof_enable(8)
        for _ in range(5):
            S.sweep_apply_T(op, prefs, W)
        gms, gn = ctx.prof_read()
        ctx.prof_enable(0)
        r[f"apply_B{B}"] = dict(ms_with_setup=dt * 1e3, tflops_with_setup=2.0 * N * N * B / dt / 1e12,
                                gemm_kernel_ms=gms / gn, gemm_tflops=2.0 * N * N * B / (gms / gn) / 1e9)
    # a solve over a 64-column sub-lattice (the 8 corners + interior points)
    idx = np.linspace(0, 4095, 64).astype(int)
    prefs = lattice[idx]
    t0 = time.perf_counter()
    Wd, iters, errs = S.sweep_solve(op, prefs, tol=1e-7, max_iter=40000)
    ctx.sync()
    dt = time.perf_counter() - t0
    steps = int(iters.max())
    r["solve_B64"] = dict(seconds=dt, max_iters=steps, min_iters=int(iters.min()),
                          ms_per_step=dt / steps * 1e3, tflops=2.0 * N * N * 64 * steps / dt / 1e12)
    for B in (512, 4096):
        t0 = time.perf_counter()
        Wd2, it2, er2, info = S.sweep_solve(op, lattice[:B], algorithm="newton", return_info=True)
        ctx.sync()
        dt2 = time.perf_counter() - t0
        r[f"newton_solve_B{B}"] = dict(seconds=dt2, gemms=int(info["gemms"]), outer_min=int(it2.min()), outer_max=int(it2.max()),
                                       inner_total_max=int(info["inner_total"].max()),
                                       tflops=2.0 * N * N * B * info["gemms"] / dt2 / 1e12,
                                       any_nan=bool(np.isnan(np.asarray(Wd2)).any()))
    # spot-check two columns against single-column device solves
    for j in (0, 63):
        m = S.SSY(γ=prefs[j, 0], ψ=prefs[j, 1], β=prefs[j, 2])
        op1 = S.make_T_ssy(m, shapes, storage="dense")
        w1, k1 = S.successive_approx(op1, np.full(shapes, 800.0), verbose=False)
        r[f"check_col{j}"] = dict(iters_sweep=int(iters[j]), iters_single=int(k1),
                                  max_rel_diff=float(np.max(np.abs(np.asarray(Wd)[j] - np.asarray(w1)) / np.asarray(w1))))
    out["C5 sweep SSY (10,)^4"] = r
    print(json.dumps({"C5": r}), flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
