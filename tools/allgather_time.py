"""Cost of the one exchange a slab-sharded factor-form apply would need: in-place all-gather of an
N-vector (8N bytes in total) across the ranks, NCCL over NVLink.  Run under torchrun."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
import sdfs_via_autodiff_b200 as S
from sdfs_via_autodiff_b200 import dist as sd
from sdfs_via_autodiff_b200._lib import lib, check
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", init_method="env://")
ctx = S.Context(local)
sd.init_comm(ctx, rank, world, dist, max_N=1024)
for N in (1048576, 9834496, 16777216):
    per = N // world
    buf = ctx.full((per * world,), float(rank))
    for _ in range(5):
        check(lib.sdfs_comm_allgather_f64(ctx.handle, buf.ptr, per), ctx.handle)
    ctx.sync(); dist.barrier()
    ctx.timer_start()
    reps = 50
    for _ in range(reps):
        check(lib.sdfs_comm_allgather_f64(ctx.handle, buf.ptr, per), ctx.handle)
    ms = ctx.timer_stop_ms() / reps
    if rank == 0:
        print(f"all-gather of {8 * per * world / 1e6:.1f} MB over {world} ranks: {ms:.3f} ms "
              f"({8 * per * (world - 1) / ms / 1e6:.0f} GB/s received per rank)", flush=True)
dist.barrier()
