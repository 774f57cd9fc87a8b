"""Row-sharded NeuMF tables (BASELINE.json configs[3]).

One GPU: the sharded addressing of the fused kernels (row r -> shard r % G, local row r // G, gradient REDs
and touched bits into the owner's shard) is checked with all G shards living in one process against
the unsharded model -- same kernels, same row-sparse (lazy) Adam, so results agree to atomic-order noise.
Two GPUs (`gpurun --gpus 2`): real NVLink peer memory + cross-GPU barriers, and the NCCL all-to-all
baseline, both against the single-process unsharded model on the global batch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _full_init(net, E, hidden):
    n_dense = int(sum(np.prod(net._offsets[k][1]) for k in net.DENSE_ORDER))
    d = {n: getattr(net, n).w.cpu().numpy().copy() for n in ("uMLP", "iMLP", "uMF", "iMF")}
    d["dense"] = net.dense.w.cpu().numpy().reshape(-1)[:n_dense].copy()
    return d


def _batches(U, I, B, steps, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(steps):
        u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
        y = (rng.random(B) < 0.25).astype(np.float32)
        out.append((u, i, y))
    return out


def test_neumf_marks_touched_rows_and_lazy_adam_moves_only_them(dev):
    from binrec_b200.NeuMFModel import NeuMFNet
    U, I, E, B = 301, 203, 8, 257
    net = NeuMFNet(U, I, E, dropout=0.0, sparse_adam="lazy", device=dev)
    w0 = {n: getattr(net, n).w.clone() for n in ("uMLP", "iMLP", "uMF", "iMF")}
    (u, i, y), = _batches(U, I, B, 1, 0)
    net.forward_backward(*(torch.from_numpy(x).to(dev) for x in (u, i, y)))
    for name, ids, rows in (("uMLP", u, U), ("uMF", u, U), ("iMLP", i, I), ("iMF", i, I)):
        bits = np.unpackbits(getattr(net, name).touched.cpu().numpy().view(np.uint8), bitorder="little")[:rows]
        expect = np.zeros(rows, dtype=np.uint8); expect[np.unique(ids)] = 1
        assert np.array_equal(bits, expect), name
    net.optimizer.apply(net.tables(), dense=[net.dense])
    for name, ids, rows in (("uMLP", u, U), ("iMF", i, I)):
        moved = (getattr(net, name).w != w0[name]).any(dim=1).cpu().numpy()
        hit = np.zeros(rows, dtype=bool); hit[np.unique(ids)] = True
        assert not moved[~hit].any(), name                     # untouched rows did not move
        assert moved[hit].mean() > 0.9, name
        assert int(getattr(net, name).touched.abs().sum().item()) == 0


@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("E", [8, 64])
def test_sharded_addressing_equals_unsharded_one_gpu(dev, G, E):
    from binrec_b200.NeuMFModel import NeuMFNet
    from binrec_b200.sharded import ShardedNeuMFNet
    U, I, B = 301, 203, 500                                  # row counts not divisible by G
    hidden = (E, E // 2, E // 4)
    ref = NeuMFNet(U, I, E, dropout=0.2, sparse_adam="lazy", device=dev)
    sh = ShardedNeuMFNet(U, I, E, dropout=0.2, device=dev, mode="peer", emulate=G, full_init=_full_init(ref, E, hidden))
    for step, (u, i, y) in enumerate(_batches(U, I, B, 3, E + G)):
        ud, idd, yd = (torch.from_numpy(x).to(dev) for x in (u, i, y))
        l0, o0 = ref.train_on_batch(ud, idd, yd, first_index=step * B, epoch=1)
        l1, o1 = sh.train_on_batch(ud, idd, yd, first_index=step * B, epoch=1)
        np.testing.assert_allclose(o1.cpu().numpy(), o0.cpu().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(l1.item(), l0.item(), rtol=1e-5, atol=1e-6)
    for name in ("uMLP", "iMLP", "uMF", "iMF"):
        np.testing.assert_allclose(getattr(sh, name).full_weights(), getattr(ref, name).w.cpu().numpy(), rtol=1e-4,
                                   atol=2e-6, err_msg=name)
    n = sh.dense.w.numel()
    np.testing.assert_allclose(sh.dense.w.cpu().numpy().reshape(-1), ref.dense.w.cpu().numpy().reshape(-1)[:n], rtol=1e-4,
                               atol=2e-6)


@pytest.mark.parametrize("G", [2, 3])
def test_sharded_bpr_equals_unsharded_one_gpu(dev, G):
    from binrec_b200.BPRModel import BPRNet
    from binrec_b200.sharded import ShardedBPRNet
    U, I, d, B = 301, 203, 64, 400
    ref = BPRNet(U, I, d, sparse_adam="lazy", device=dev)
    sh = ShardedBPRNet(U, I, d, device=dev, emulate=G, full_init={"user": ref.user.w.cpu().numpy(), "item": ref.item.w.cpu().numpy()})
    rng = np.random.default_rng(G)
    for step in range(3):
        u, p, n = ((U * rng.random(B) ** 2).astype(np.int32), (I * rng.random(B) ** 2).astype(np.int32),
                   rng.integers(0, I, B).astype(np.int32))
        ud, pd_, nd = (torch.from_numpy(x).to(dev) for x in (u, p, n))
        l0 = ref.train_on_batch(ud, pd_, nd)
        l1 = sh.train_on_batch(ud, pd_, nd)
        np.testing.assert_allclose(l1.item(), l0.item(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(sh.user.full_weights(), ref.user.w.cpu().numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(sh.item.full_weights(), ref.item.w.cpu().numpy(), rtol=1e-5, atol=2e-6)


# ---- two GPUs ------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


U2, I2, E2, B2, STEPS2 = 1001, 403, 16, 512, 3


def _worker(rank, world, port, mode, init, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", device_id=dev)
    try:
        from binrec_b200 import distributed as D
        from binrec_b200.sharded import ShardedNeuMFNet
        net = ShardedNeuMFNet(U2, I2, E2, dropout=0.0, device=dev, mode=mode, full_init=init)
        losses = []
        for step, (u, i, y) in enumerate(_batches(U2, I2, world * B2, STEPS2, 5)):
            lo, hi = D.local_slice(world * B2)
            l, _ = net.train_on_batch(*(torch.from_numpy(x[lo:hi]).to(dev) for x in (u, i, y)), first_index=step * world * B2 + lo)
            losses.append(float(l.item()))
        net.check()
        full = {n: getattr(net, n).full_weights() for n in ("uMLP", "iMLP", "uMF", "iMF")}
        bpr = {}
        if mode == "peer":                    # row-sharded BPR on the same two GPUs
            from binrec_b200.sharded import ShardedBPRNet
            b = ShardedBPRNet(U2, I2, 64, device=dev, full_init=init["bpr"])
            rngb = np.random.default_rng(21)
            for step in range(3):
                u, p, n = (rngb.integers(0, U2, world * B2).astype(np.int32), rngb.integers(0, I2, world * B2).astype(np.int32),
                           rngb.integers(0, I2, world * B2).astype(np.int32))
                lo, hi = D.local_slice(world * B2)
                b.train_on_batch(*(torch.from_numpy(x[lo:hi]).to(dev) for x in (u, p, n)))
            b.check()
            bpr = dict(bpr_user=b.user.full_weights(), bpr_item=b.item.full_weights())
        ret[(mode, rank)] = dict(losses=losses, dense=net.dense.w.cpu().numpy().reshape(-1), **full, **bpr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_sharded_neumf_two_gpus_matches_unsharded(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from binrec_b200.NeuMFModel import NeuMFNet
    world = 2
    dev = torch.device("cuda:0")
    hidden = (E2, E2 // 2, E2 // 4)
    # BatchNorm statistics are per replica in the sharded run (MirroredStrategy's default), so the single-process
    # reference runs the two half batches as two forward/backward passes scaled by 1/(global batch), then one
    # optimizer step -- exactly what the two ranks do together
    ref = NeuMFNet(U2, I2, E2, dropout=0.0, sparse_adam="lazy", device=dev)
    init = _full_init(ref, E2, hidden)
    from binrec_b200.BPRModel import BPRNet
    bref = BPRNet(U2, I2, 64, sparse_adam="lazy", device=dev)
    init["bpr"] = {"user": bref.user.w.cpu().numpy().copy(), "item": bref.item.w.cpu().numpy().copy()}
    rngb = np.random.default_rng(21)
    for step in range(3):                    # the global batch in one process (BPR has no per-replica statistics)
        u, p, n = (rngb.integers(0, U2, world * B2).astype(np.int32), rngb.integers(0, I2, world * B2).astype(np.int32),
                   rngb.integers(0, I2, world * B2).astype(np.int32))
        bref.train_on_batch(*(torch.from_numpy(x).to(dev) for x in (u, p, n)))
    ref_losses = []
    for step, (u, i, y) in enumerate(_batches(U2, I2, world * B2, STEPS2, 5)):
        ls = []
        for r in range(world):
            sl = slice(r * B2, (r + 1) * B2)
            l, _ = ref.forward_backward(*(torch.from_numpy(x[sl]).to(dev) for x in (u, i, y)),
                                        first_index=step * world * B2 + r * B2, global_batch=world * B2)
            ls.append(float(l.item()))
            if r == 0:
                bn_after_first = ref.bn_moving.clone()
        ref.optimizer.apply(ref.tables(), dense=[ref.dense])
        ref_losses.append(ls)
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, init, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    for r in range(world):
        got = ret[(mode, r)]
        np.testing.assert_allclose(got["losses"], [ls[r] for ls in ref_losses], rtol=1e-5, atol=1e-6)
        for name in ("uMLP", "iMLP", "uMF", "iMF"):
            np.testing.assert_allclose(got[name], getattr(ref, name).w.cpu().numpy(), rtol=1e-4, atol=2e-6, err_msg=name)
        n = len(got["dense"])
        np.testing.assert_allclose(got["dense"], ref.dense.w.cpu().numpy().reshape(-1)[:n], rtol=1e-4, atol=2e-6)
    assert np.array_equal(ret[(mode, 0)]["dense"], ret[(mode, 1)]["dense"])
    if mode == "peer":
        np.testing.assert_allclose(ret[(mode, 0)]["bpr_user"], bref.user.w.cpu().numpy(), rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(ret[(mode, 0)]["bpr_item"], bref.item.w.cpu().numpy(), rtol=1e-5, atol=2e-6)


def test_standalone_sharded_gather_and_scatter_add(dev):
    """brk_gather_rows_sharded / brk_scatter_add_rows_sharded (the all-to-all of rows and of row gradients as ops) with
    the G shards of one table held by one process: bit-exact gather, exact integer scatter-add into the owners'
    accumulators, owners' touched bits == the distinct ids."""
    from binrec_b200 import hotpath as H
    from binrec_b200.sharded import ShardedTable
    rng = np.random.default_rng(0)
    for G, rows, d in ((3, 1000, 64), (8, 4097, 32), (2, 77, 8)):
        full = rng.integers(-8, 9, size=(rows, d)).astype(np.float32)
        tab = ShardedTable(rows, d, G, 0, dev, full_init=full, symmetric=False, emulate=True)
        ids = torch.from_numpy((rows * rng.random(5000) ** 2).astype(np.int32)).to(dev)
        out = H.gather_rows_sharded(tab.c_shards(), d, ids)
        assert np.array_equal(out.cpu().numpy(), full[ids.cpu().numpy()])
        vals = torch.from_numpy(rng.integers(-4, 5, size=(5000, d)).astype(np.float32)).to(dev)
        H.scatter_add_rows_sharded(tab.c_shards(), d, ids, vals)
        ref = np.zeros((rows, d), np.float32)
        np.add.at(ref, ids.cpu().numpy(), vals.cpu().numpy())
        for r in range(G):
            mine = ref[r::G]
            assert np.array_equal(tab.tables[r].g.cpu().numpy()[:len(mine)], mine)
            bits = np.zeros(tab.local_rows, bool); loc = np.unique(ids.cpu().numpy()[ids.cpu().numpy() % G == r] // G); bits[loc] = True
            words = tab.tables[r].touched.cpu().numpy().view(np.uint32)
            got = ((words[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(bool).reshape(-1)[:tab.local_rows]
            assert np.array_equal(got, bits)
