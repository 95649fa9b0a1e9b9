"""BASELINE config 5 as named: the 16^3 (gamma, psi, beta) lattice = 4096 parameter sets of SSY (10,)^4,
columns sharded over the ranks (512 per GPU at 8 ranks), P replicated; Newton and SA sweeps.
Run under torchrun, one rank per GPU.  Wall clock from a barrier before the solve to the gathered result."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
import sdfs_via_autodiff_b200 as S
from sdfs_via_autodiff_b200 import dist as sd
from sdfs_via_autodiff_b200.dist import TorchExchange

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", init_method="env://")
ctx = S.Context(local)
S.Context._default = ctx
sd.init_comm(ctx, rank, world, dist, max_N=120000)
shapes = (10,) * 4
form = "dense" if "--dense" in sys.argv else "factor"
op = S.make_sweep_operator(S.SSY(), shapes, ctx=ctx, form=form)
g = np.linspace(5, 12, 16); p = np.linspace(1.3, 2.0, 16); b = np.linspace(0.997, 0.999, 16)
lattice = np.array([[gi, pi, bi] for gi in g for pi in p for bi in b])
ex = TorchExchange(dist)
out = {"world": world, "sets": len(lattice), "shapes": shapes, "form": form}
for algo, kw in (("newton", {}), ("successive_approx", {})):
    if algo == "successive_approx" and "--sa" not in sys.argv:
        continue
    S.sweep_solve(op, lattice[: 8 * world], algorithm=algo, exchange=ex, **kw)      # warm-up (allocations, NCCL)
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    W, it, er = S.sweep_solve(op, lattice, algorithm=algo, exchange=ex, **kw)
    ctx.sync(); dist.barrier()
    dt = time.perf_counter() - t0
    out[algo] = dict(seconds=dt, iters_min=int(np.min(it)), iters_max=int(np.max(it)), nan=bool(np.isnan(W).any()),
                     w_min=float(np.min(W)), w_max=float(np.max(W)), sets_per_s=len(lattice) / dt)
    if algo == "newton" and rank == 0:
        # corners against single-column solves on this rank's GPU
        chk = []
        for j in (0, 15, 255, 4095):
            m = S.SSY(γ=lattice[j, 0], ψ=lattice[j, 1], β=lattice[j, 2])
            op1 = S.make_T_ssy(m, shapes, storage="kron", ctx=ctx)
            w1, k1 = S.newton_solver(op1, np.full(shapes, 800.0), verbose=False)
            chk.append(dict(col=j, outer_sweep=int(it[j]), outer_single=int(k1),
                            max_rel_diff=float(np.max(np.abs(W[j] - np.asarray(w1)) / np.asarray(w1)))))
        out["corner_check"] = chk
if rank == 0:
    print(json.dumps(out), flush=True)
dist.barrier()
