"""Host-side multi-rank logic on CPU with the gloo backend, world_size 2 (no GPU needed): the
row-shard routing (ids -> owners -> rows -> batch order, and gradients back), the data-parallel
gradient scaling identity and the sharded top-K merge.  The device kernels are replaced here by NumPy
indexing from the oracle -- what is under test is the index arithmetic and the collective sequence."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run(rank, world, port, fn, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_run, args=(r, world, port, fn, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return dict(ret)


def _routing_case(rank, world):
    from binrec_b200 import distributed as D
    rows, d = 101, 4
    full = torch.arange(rows * d, dtype=torch.float32).view(rows, d)
    shard = full[rank::world].clone()                               # owner = id mod G, local row = id div G
    assert shard.shape[0] == D.shard_rows(rows, rank, world)
    g = torch.Generator().manual_seed(100 + rank)
    ids = torch.randint(0, rows, (37 + 5 * rank,), generator=g)
    lk = D.ShardedLookup(ids, world)
    got = lk.forward(lambda local: shard[local])
    ok_fwd = torch.equal(got, full[ids])
    grad_shard = torch.zeros_like(shard)
    lk.backward(torch.ones(len(ids), d) * (rank + 1), lambda local, vals: grad_shard.index_add_(0, local, vals))
    # expected: every rank's ids contribute (rank+1) to the owner's row
    all_ids = [torch.randint(0, rows, (37 + 5 * r,), generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
    exp = torch.zeros(rows, d)
    for r, a in enumerate(all_ids):
        exp.index_add_(0, a, torch.ones(len(a), d) * (r + 1))
    return bool(ok_fwd and torch.equal(grad_shard, exp[rank::world]))


def test_row_shard_routing_world2():
    assert all(_spawn(_routing_case).values())


def _dp_case(rank, world):
    """Sum over ranks of gradients scaled by 1/(world*B) == gradient of the global-batch mean loss."""
    from binrec_b200 import distributed as D
    from oracle import bpr as OB
    rng = np.random.default_rng(0)
    U, I, d, B = 40, 30, 8, 16
    user = rng.normal(size=(U, d)); item = rng.normal(size=(I, d))
    u = rng.integers(0, U, world * B); p = rng.integers(0, I, world * B); n = rng.integers(0, I, world * B)
    lo, hi = D.local_slice(world * B)
    assert (lo, hi) == (rank * B, (rank + 1) * B)
    _, gu, gi = OB.bpr_loss_and_grads(user, item, u[lo:hi], p[lo:hi], n[lo:hi])
    flat = torch.from_numpy(np.concatenate([gu.ravel(), gi.ravel()]) / world)     # local mean -> 1/(world*B)
    D.all_reduce_sum_(flat)
    _, Gu, Gi = OB.bpr_loss_and_grads(user, item, u, p, n)
    return bool(np.allclose(flat.numpy(), np.concatenate([Gu.ravel(), Gi.ravel()]), rtol=1e-12, atol=1e-15))


def test_data_parallel_gradient_identity_world2():
    assert all(_spawn(_dp_case).values())


def _topk_case(rank, world):
    from binrec_b200 import distributed as D
    from oracle import topk as OT
    rng = np.random.default_rng(1)
    Q = (rng.integers(-4, 5, size=(19, 8)) / 8.0).astype(np.float32)
    C = (rng.integers(-4, 5, size=(64, 8)) / 8.0).astype(np.float32)
    lo, hi = D.local_slice(64)

    def local_topk(q, c, off):
        v, i = OT.brute_force_topk(q.numpy(), c.numpy(), 5)
        return torch.from_numpy(v), torch.from_numpy(i + off)

    def merge(pv, pi):
        v, i = OT.merge_topk(list(pv.numpy()), list(pi.numpy()), 5)
        return torch.from_numpy(v), torch.from_numpy(i)

    v, i = D.sharded_topk(torch.from_numpy(Q), torch.from_numpy(C[lo:hi]), lo, 5, merge=merge, local_topk=local_topk)
    rv, ri = OT.brute_force_topk(Q, C, 5)
    return bool(np.array_equal(i.numpy(), ri) and np.array_equal(v.numpy(), rv))


def test_sharded_topk_merge_world2():
    assert all(_spawn(_topk_case).values())


def test_bucket_by_owner_is_stable_partition():
    from binrec_b200 import distributed as D
    ids = torch.tensor([5, 2, 9, 4, 7, 2, 8, 1])
    perm, counts = D.bucket_by_owner(ids, 3)
    assert counts.tolist() == [1, 3, 4]
    assert ids[perm].tolist() == [9, 4, 7, 1, 5, 2, 2, 8]
    assert D.local_row(torch.tensor([0, 1, 2, 3, 7]), 3).tolist() == [0, 0, 0, 1, 2]
    assert [D.shard_rows(10, r, 3) for r in range(3)] == [4, 3, 3]
    assert [D.local_slice(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
