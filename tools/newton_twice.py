"""First-call vs steady-state time of the fused multi-GPU Newton loop (run under torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
import sdfs_via_autodiff_b200 as S
from sdfs_via_autodiff_b200 import dist as sd
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo", init_method="env://")
ctx = S.Context(local)
S.Context._default = ctx
if world > 1:
    sd.init_comm(ctx, rank, world, dist, max_N=120000)
shapes = (18,) * 4
t0 = time.perf_counter(); op = S.make_T_ssy(S.SSY(), shapes, storage="dense", ctx=ctx); ctx.sync()
if rank == 0: print("build %.3f s" % (time.perf_counter() - t0), flush=True)
w0 = ctx.full(shapes, 800.0)
for i in range(3):
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    w, k, info = S.newton_solver(op, w0, tol=1e-8, verbose=False, return_info=True)
    ctx.sync()
    if rank == 0: print("newton call %d: %.3f s, %d applications, %.3f ms/app" % (i, time.perf_counter() - t0, info["matvecs"], (time.perf_counter() - t0) / info["matvecs"] * 1e3), flush=True)
for i in range(2):
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    w, k = S.successive_approx(op, w0, tol=1e-8, max_iter=200, verbose=False)
    ctx.sync()
    if rank == 0: print("SA 200 its call %d: %.3f s" % (i, time.perf_counter() - t0), flush=True)
