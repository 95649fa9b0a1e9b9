"""Multi-GPU parity check (run under torchrun, one rank per GPU):
row-sharded dense operator vs the oracle; NCCL all-gather path for single applications,
fused peer-store path for the device-resident SA / Newton loops."""
import os, sys, time
os.environ.setdefault("SDFS_KRON_SHARD_MIN", "0")      # the small test grids are slab-sharded too (default: >= 4 M states)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
import oracle as O
import sdfs_via_autodiff_b200 as S
from sdfs_via_autodiff_b200 import dist as sd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", init_method="env://")
ctx = S.Context(local)
S.Context._default = ctx
sd.init_comm(ctx, rank, world, dist, max_N=120000)
ok = True


def report(name, cond, extra=""):
    global ok
    ok = ok and bool(cond)
    if rank == 0:
        print(("PASS " if cond else "FAIL ") + name + " " + extra, flush=True)


ssy = O.SSY()
for shapes in ((2, 3, 4, 5), (7, 5, 6, 9), (10, 10, 10, 10)):
    arrays = O.discretize_ssy(ssy, shapes)
    kop = O.KronSSY(shapes, ssy.params, arrays)
    op = S.make_T_ssy(S.SSY(), shapes, storage="dense", ctx=ctx)
    N = op.N
    report(f"{shapes} row slice", (op.row_begin, op.row_end) == sd.row_partition(N, world, rank),
           f"rank0 rows [{op.row_begin},{op.row_end})")
    rng = np.random.default_rng(1233)
    w = np.exp(rng.standard_normal(shapes))
    got = np.asarray(op(w))
    path = "NCCL all-gather" if os.environ.get("SDFS_FUSED_EXCHANGE") == "0" else "fused peer-store exchange"
    report(f"{shapes} T ({path})", np.allclose(got, kop.T(w), rtol=1e-12, atol=0))
    # chained applications without host round trips: exercises the double-buffered result slots
    wd = ctx.asarray(w) if hasattr(ctx, "asarray") else op._in(w)
    ref = w
    for _ in range(7):
        wd = op(wd)
        ref = kop.T(ref)
    report(f"{shapes} 7 chained T ({path})", np.allclose(np.asarray(wd), ref, rtol=1e-11, atol=0))
    report(f"{shapes} P 1 = 1 ({path})", np.allclose(np.asarray(op.apply_P(np.ones(shapes))), 1.0, rtol=0, atol=1e-12))
    v = rng.standard_normal(shapes)
    report(f"{shapes} JVP", np.allclose(np.asarray(op.jvp(w, v)), kop.jvp(w, v), rtol=1e-10, atol=1e-12))
    t0 = time.time()
    ws, k = S.successive_approx(op, np.full(shapes, 800.0), verbose=False)
    dt = time.time() - t0
    w_ref, k_ref = O.successive_approx(kop.T, np.full(shapes, 800.0), verbose=False)
    report(f"{shapes} SA fused loop", abs(k - k_ref) <= 1 and np.allclose(np.asarray(ws), w_ref, rtol=1e-10),
           f"iters {k} vs {k_ref}, {dt:.2f}s, {dt / max(k, 1) * 1e6:.1f} us/iter")
    t0 = time.time()
    wn, kn, info = S.newton_solver(op, np.full(shapes, 800.0), verbose=False, return_info=True)
    dt = time.time() - t0
    wn_ref, kn_ref = O.newton_solver(kop.T, np.full(shapes, 800.0), jvp=kop.jvp, verbose=False)
    report(f"{shapes} Newton fused loop", abs(kn - kn_ref) <= 1 and np.allclose(np.asarray(wn), wn_ref, rtol=1e-5),
           f"outer {kn} vs {kn_ref}, inner {info['inner_iters']}, {dt:.2f}s")
    wt, kt = S.newton_solver(op, np.full(shapes, 800.0), tol=1e-9, bicgstab_atol=1e-10, krylov_rtol=1e-12, verbose=False)
    wfix, _ = O.successive_approx(kop.T, wn_ref, tol=1e-11, verbose=False)
    report(f"{shapes} Newton tight", np.allclose(np.asarray(wt), wfix, rtol=1e-10))
    wg, kg = S.newton_solver(op, np.full(shapes, 800.0), krylov="gmres", tol=1e-9, bicgstab_atol=1e-10,
                             krylov_rtol=1e-12, verbose=False)
    report(f"{shapes} Newton GMRES", np.allclose(np.asarray(wg), wfix, rtol=1e-10))
    q, e = op.sdf(wt)
    report(f"{shapes} SDF euler", np.max(np.abs(np.asarray(e))) < 1e-8)
    del op
# factor-form operator, leading axis split into per-rank slabs (h_lambda contraction first, result rows exchanged
# by peer stores); (11, ...) gives ragged slabs at 2 and 4 ranks, (9, ...) leaves ranks without rows at 8 ranks
for shapes in ((12, 5, 10, 11), (11, 4, 9, 10), (9, 3, 18, 17)):
    arrays = O.discretize_ssy(ssy, shapes)
    kop = O.KronSSY(shapes, ssy.params, arrays)
    op = S.make_T_ssy(S.SSY(), shapes, storage="kron", ctx=ctx)
    N = op.N
    want = sd.slab_partition(shapes, world, rank)
    report(f"{shapes} kron slab rows", (op.row_begin, op.row_end) == want and (world == 1 or op.row_end - op.row_begin < N),
           f"rank0 rows [{op.row_begin},{op.row_end})")
    rng = np.random.default_rng(1233)
    w = np.exp(rng.standard_normal(shapes))
    report(f"{shapes} kron T", np.allclose(np.asarray(op(w)), kop.T(w), rtol=1e-12, atol=0))
    wd = ctx.asarray(w)
    ref = w
    for _ in range(7):
        wd = op(wd)
        ref = kop.T(ref)
    report(f"{shapes} kron 7 chained T", np.allclose(np.asarray(wd), ref, rtol=1e-11, atol=0))
    report(f"{shapes} kron P 1 = 1", np.allclose(np.asarray(op.apply_P(np.ones(shapes))), 1.0, rtol=0, atol=1e-12))
    v = rng.standard_normal(shapes)
    report(f"{shapes} kron JVP", np.allclose(np.asarray(op.jvp(w, v)), kop.jvp(w, v), rtol=1e-10, atol=1e-12))
    ws, k = S.successive_approx(op, np.full(shapes, 800.0), tol=1e-6, verbose=False)
    w_ref, k_ref = O.successive_approx(kop.T, np.full(shapes, 800.0), tol=1e-6, verbose=False)
    report(f"{shapes} kron SA fused loop", abs(k - k_ref) <= 1 and np.allclose(np.asarray(ws), w_ref, rtol=1e-10),
           f"iters {k} vs {k_ref}")
    wn, kn, info = S.newton_solver(op, np.full(shapes, 800.0), verbose=False, return_info=True)
    wn_ref, kn_ref = O.newton_solver(kop.T, np.full(shapes, 800.0), jvp=kop.jvp, verbose=False)
    report(f"{shapes} kron Newton fused loop", abs(kn - kn_ref) <= 1 and np.allclose(np.asarray(wn), wn_ref, rtol=1e-5),
           f"outer {kn} vs {kn_ref}, inner {info['inner_iters']}")
    wt, kt = S.newton_solver(op, np.full(shapes, 800.0), tol=1e-9, bicgstab_atol=1e-10, krylov_rtol=1e-12, verbose=False)
    wfix, _ = O.successive_approx(kop.T, wn_ref, tol=1e-11, verbose=False)
    report(f"{shapes} kron Newton tight", np.allclose(np.asarray(wt), wfix, rtol=1e-10))
    wg, kg = S.newton_solver(op, np.full(shapes, 800.0), krylov="gmres", tol=1e-9, bicgstab_atol=1e-10,
                             krylov_rtol=1e-12, verbose=False)
    report(f"{shapes} kron Newton GMRES", np.allclose(np.asarray(wg), wfix, rtol=1e-10))
    wa, ka = S.solvers["anderson"](op, np.full(shapes, 800.0), tol=1e-6, verbose=False)
    report(f"{shapes} kron Anderson", np.allclose(np.asarray(wa), wfix, rtol=1e-6), f"iters {ka}")
    q, e = op.sdf(wt)
    from oracle.sdf import e_sdf_ssy
    P_, ar_, ac_, β_, θ_ = O.dense_ssy(shapes, ssy.params, arrays)
    q_ref, _ = O.sdf_dense(np.asarray(wt), P_, ar_, ac_, e_sdf_ssy(shapes, ssy.params, arrays), β_, θ_)
    report(f"{shapes} kron SDF", np.max(np.abs(np.asarray(e))) < 1e-8 and np.allclose(np.asarray(q).reshape(-1), q_ref, rtol=1e-10))
    # bit-identical to the same operator kept whole on one rank (same contraction order, same arithmetic per row)
    whole = S.WCOperator.from_factors(S.Factors.build(0, S.SSY().params, shapes, ctx), storage="kron_local")
    report(f"{shapes} kron sharded == whole, bit for bit", np.array_equal(np.asarray(op(w)), np.asarray(whole(w))))
    del op, whole
# GCY factor form stays whole on every rank (its leading axis z is contracted last): still correct in a multi-rank context
gcy = O.GCY()
gshapes = (3, 4, 3, 2, 3, 4)
garr = O.discretize_gcy(gcy, gshapes)
gk = O.KronGCY(gshapes, gcy.params, garr)
gop = S.make_T_gcy(S.GCY(), gshapes, storage="kron", ctx=ctx)
wg_ = np.exp(np.random.default_rng(7).standard_normal(gshapes))
report("GCY kron (rank-local) T", (gop.row_begin, gop.row_end) == (0, gop.N) and np.allclose(np.asarray(gop(wg_)), gk.T(wg_), rtol=1e-12))
del gop
# parameter sweep: columns sharded over the ranks, P replicated, results gathered on the host
from sdfs_via_autodiff_b200.dist import TorchExchange
shapes = (4, 7, 6, 5)
arrays = O.discretize_ssy(ssy, shapes)
prefs = np.array([[8.89, 1.97, 0.999], [5.0, 1.3, 0.997], [12.0, 2.0, 0.999], [7.3, 1.61, 0.998], [10.0, 1.5, 0.9985]])
for form in ("dense", "factor"):
    sop = S.make_sweep_operator(S.SSY(), shapes, ctx=ctx, form=form)
    Wall, its, errs = S.sweep_solve(sop, prefs, algorithm="newton", tol=1e-9, bicgstab_atol=1e-10, krylov_rtol=1e-12,
                                    exchange=TorchExchange(dist))
    good = Wall.shape == (len(prefs),) + shapes
    for b, (γ, ψ, β) in enumerate(prefs):
        m = O.SSY(γ=γ, ψ=ψ, β=β)
        kop = O.KronSSY(shapes, m.params, arrays)
        w_ref, _ = O.newton_solver(kop.T, np.full(shapes, 800.0), jvp=kop.jvp, bicgstab_atol=1e-11, verbose=False)
        w_ref, _ = O.successive_approx(kop.T, w_ref, tol=1e-11, verbose=False)
        good = good and np.allclose(Wall[b], w_ref, rtol=1e-10)
    report(f"sweep[{form}] (newton) sharded over columns", good, f"outer iters {list(its)}")
    Wsa, its_sa, _ = S.sweep_solve(sop, prefs[:3], algorithm="successive_approx", tol=1e-6, exchange=TorchExchange(dist))
    report(f"sweep[{form}] (SA) sharded over columns", Wsa.shape[0] == 3 and np.isfinite(Wsa).all(), f"iters {list(its_sa)}")
flag = [ok]
allok = [None] * world
dist.all_gather_object(allok, ok)
if rank == 0:
    print("ALL PASS" if all(allok) else f"SOME FAILED {allok}", flush=True)
dist.barrier()
dist.destroy_process_group()
