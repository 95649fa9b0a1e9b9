"""World-size-2 gloo tests (CPU): the host-side logic of the multi-GPU path -- row
partition, id/handle exchange, and the shard -> all-gather structure of one operator
application, emulated with the oracle on each rank."""
import os
import socket

import numpy as np
import pytest

import oracle as O
from sdfs_via_autodiff_b200.dist import row_partition, slab_partition, TorchExchange


def test_row_partition_covers_all_rows():
    for N in (1, 7, 120, 10000, 104976, 117649):
        for G in (1, 2, 3, 4, 8):
            parts = [row_partition(N, G, r) for r in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == N
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1 and b0 <= e0
            assert max(e - b for b, e in parts) == (N + G - 1) // G


def test_slab_partition_covers_the_leading_axis():
    for shapes in ((12, 5, 10, 11), (11, 4, 9, 10), (9, 3, 18, 17), (56, 56, 56, 56)):
        N = int(np.prod(shapes))
        for G in (1, 2, 3, 4, 8):
            parts = [slab_partition(shapes, G, r) for r in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == N
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1 and b0 <= e0
            inner = N // shapes[0]
            assert all(b % inner == 0 and e % inner == 0 for b, e in parts)


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = TorchExchange(dist)
    uid = os.urandom(128) if rank == 0 else None
    uid = ex.bcast(uid, 0)
    handles = ex.allgather(bytes([rank]) * 64)
    assert len(uid) == 128 and [h[0] for h in handles] == list(range(world))
    # one row-sharded operator application: local rows, then all-gather of the slices
    ssy = O.SSY()
    shapes = (3, 4, 5, 6)
    P, ar, ac, β, θ = O.dense_ssy(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
    N = P.shape[0]
    w = np.exp(np.random.default_rng(1233).standard_normal(N))
    b, e = row_partition(N, world, rank)
    local = 1 + β * (ar[b:e] * (P[b:e] @ (ac * w ** θ))) ** (1 / θ)
    full = np.concatenate(ex.allgather(local))
    np.testing.assert_allclose(full, O.dense_T(w, P, ar, ac, β, θ), rtol=1e-14)
    # one slab-sharded factor-form application: the leading-axis contraction first, restricted to the own
    # slab's rows (it alone reads the other ranks' part of the input), the rest local, then the gather
    kop = O.KronSSY(shapes, ssy.params, O.discretize_ssy(ssy, shapes))
    sb, se = slab_partition(shapes, world, rank)
    inner = N // shapes[0]
    l0, l1 = sb // inner, se // inner
    x = (kop.a_col * w.reshape(shapes) ** kop.θ)
    U = np.einsum('ab,bkij->akij', kop.Q_λ[l0:l1], x)           # rows l0..l1 only, full input
    U = np.einsum('ab,lkbj->lkaj', kop.Q_hz, U)
    U = np.einsum('ijq,lkiq->lkij', kop.z_Q, U)
    U = np.einsum('ab,lbij->laij', kop.Q_c, U)
    loc = 1 + kop.β * (kop.a_row[l0:l1] * U) ** (1 / kop.θ)
    full_k = np.concatenate([p.reshape(-1) for p in ex.allgather(loc)])
    np.testing.assert_allclose(full_k, kop.T(w.reshape(shapes)).reshape(-1), rtol=1e-13)
    # replicated stopping decision: every rank reduces the same full vectors
    err = np.max(np.abs(full - w))
    errs = ex.allgather(float(err))
    assert len(set(errs)) == 1
    np.save(os.path.join(out_dir, f"uid{rank}.npy"), np.frombuffer(uid, dtype=np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "uid0.npy"), np.load(tmp_path / "uid1.npy")
    assert np.array_equal(a, b)
