"""Multi-GPU bring-up: one process per GPU, P row-sharded over the ranks.

The library needs two small host-side exchanges at start-up -- the 128-byte NCCL
unique id (rank 0 -> all) and the 64-byte CUDA-IPC handle of every rank's exchange
arena (all -> all).  Any transport works; ``TorchExchange`` adapts
``torch.distributed`` (gloo or nccl) without this package importing torch.
After ``init_comm`` every dense operator built on the context materialises only
its row slice, single applications end with an NCCL all-gather of the result
slices, and the device-resident solver loops exchange their vectors with peer
stores over NVLink inside the kernel (csrc/loops.cu).
"""
import ctypes as C

from ._lib import lib, check


def row_partition(N, nranks, rank):
    """Rows [begin, end) of rank ``rank``: contiguous chunks of ceil(N / nranks)."""
    chunk = (N + nranks - 1) // nranks
    b = min(N, chunk * rank)
    return b, min(N, b + chunk)


def slab_partition(shapes, nranks, rank):
    """Rows [begin, end) of a slab-sharded factor-form operator: the leading axis is cut into
    contiguous chunks of ceil(shapes[0] / nranks) indices (the last ranks may get fewer, or none)."""
    L = int(shapes[0])
    inner = 1
    for s in shapes[1:]:
        inner *= int(s)
    chunk = (L + nranks - 1) // nranks
    l0 = min(L, chunk * rank)
    return l0 * inner, min(L, l0 + chunk) * inner


class TorchExchange:
    """bcast / allgather of small Python objects over a torch.distributed process group."""

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group

    def bcast(self, obj, src=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def allgather(self, obj):
        out = [None] * self.dist.get_world_size(self.group)
        self.dist.all_gather_object(out, obj, group=self.group)
        return out


def init_comm(ctx, rank, nranks, exchange, max_N=0):
    """Create the NCCL communicator of ``ctx`` and, if ``max_N`` > 0, map the peers'
    exchange arenas (needed by the fused multi-GPU solver loops for grids up to max_N
    states).  ``exchange`` is a TorchExchange-like object or torch.distributed itself."""
    if not hasattr(exchange, "bcast"):
        exchange = TorchExchange(exchange)
    uid = None
    if rank == 0:
        buf = C.create_string_buffer(128)
        check(lib.sdfs_comm_unique_id(buf))
        uid = buf.raw
    uid = exchange.bcast(uid, 0)
    check(lib.sdfs_comm_init(ctx.handle, int(rank), int(nranks), uid), ctx.handle)
    ctx.rank, ctx.nranks = int(rank), int(nranks)
    if max_N > 0 and nranks > 1:
        h = C.create_string_buffer(64)
        check(lib.sdfs_comm_arena_export(ctx.handle, int(max_N), h), ctx.handle)
        handles = exchange.allgather(h.raw)
        allh = b"".join(handles)
        check(lib.sdfs_comm_arena_import(ctx.handle, allh), ctx.handle)
    return ctx


def barrier(ctx):
    check(lib.sdfs_comm_barrier(ctx.handle), ctx.handle)
