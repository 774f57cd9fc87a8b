"""2-GPU NCCL parity of the mirrored data-parallel step (one process per GPU) against the
single-process oracle on the global batch.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", device_id=dev)
    try:
        from binrec_b200.BPRModel import BPRNet
        from binrec_b200.NeuMFModel import NeuMFNet
        from binrec_b200 import distributed as D, hotpath as H
        U, I, d, B = 200, 150, 64, 256
        rng = np.random.default_rng(3)
        net = BPRNet(U, I, d, seed=42, device=dev)
        losses = []
        for step in range(3):
            u = rng.integers(0, U, world * B).astype(np.int32); p = rng.integers(0, I, world * B).astype(np.int32)
            n = rng.integers(0, I, world * B).astype(np.int32)
            lo, hi = D.local_slice(world * B)
            l = net.train_on_batch(*(torch.from_numpy(x[lo:hi]).to(dev) for x in (u, p, n)))
            losses.append(float(l.item()))
        # sharded top-K over item ranges, merged across the two GPUs
        Q = torch.from_numpy((np.random.default_rng(4).integers(-4, 5, size=(77, 64)) / 8.0).astype(np.float32)).to(dev)
        C = torch.from_numpy((np.random.default_rng(5).integers(-4, 5, size=(1000, 64)) / 8.0).astype(np.float32)).to(dev)
        lo, hi = D.local_slice(1000)
        tv, ti = D.sharded_topk(Q, C[lo:hi].contiguous(), lo, 10)
        # NeuMF mirrored step
        nm = NeuMFNet(U, I, 8, dropout=0.0, device=dev)
        rng2 = np.random.default_rng(6)
        u = rng2.integers(0, U, world * B).astype(np.int32); it = rng2.integers(0, I, world * B).astype(np.int32)
        y = (rng2.random(world * B) < 0.3).astype(np.float32)
        lo, hi = D.local_slice(world * B)
        nm.train_on_batch(*(torch.from_numpy(x[lo:hi]).to(dev) for x in (u, it, y)))
        # host-fed data-parallel steps: every rank feeds its own batches from pinned host memory, negatives drawn in-kernel
        rngp = np.random.default_rng(8)
        key = np.unique(rngp.integers(0, U * I, 3000))
        pu, pi = (key // I).astype(np.int32), (key % I).astype(np.int32)
        perm = rngp.permutation(len(pu)); pu, pi = pu[perm], pi[perm]
        net2 = BPRNet(U, I, d, seed=43, device=dev)
        net2.set_training_pairs(pu, pi)
        packed = BPRNet.pack_host_batches(pu, pi, B)
        order = [rank, 2 + rank, 4 + rank]
        hl = net2.train_steps_from_host(packed, None, order, B, 7, 2)
        torch.cuda.synchronize()
        ret[rank] = dict(user=net.user.w.cpu().numpy(), item=net.item.w.cpu().numpy(), losses=losses,
                         hf_user=net2.user.w.cpu().numpy(), hf_item=net2.item.w.cpu().numpy(), hf_losses=hl.numpy().copy(),
                         tv=tv.cpu().numpy(), ti=ti.cpu().numpy(), neumf_W1=nm.param("W1").cpu().numpy(),
                         neumf_uMLP=nm.uMLP.w.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_mirrored_bpr_neumf_and_sharded_topk_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import bpr as OB, topk as OT
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    U, I, d, B = 200, 150, 64, 256
    orc = OB.BPROracle(U, I, d, seed=42)
    rng = np.random.default_rng(3)
    for step in range(3):
        u = rng.integers(0, U, world * B).astype(np.int32); p = rng.integers(0, I, world * B).astype(np.int32)
        n = rng.integers(0, I, world * B).astype(np.int32)
        orc.step(u, p, n)                     # the global batch in one process
    for r in range(world):
        np.testing.assert_allclose(ret[r]["user"], orc.user, rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(ret[r]["item"], orc.item, rtol=1e-5, atol=2e-6)
    # host-fed DP path vs the oracle on the union of both ranks' batches, with the oracle's sampler
    from oracle import philox as OP
    rngp = np.random.default_rng(8)
    key = np.unique(rngp.integers(0, U * I, 3000))
    pu, pi = (key // I).astype(np.int32), (key % I).astype(np.int32)
    perm = rngp.permutation(len(pu)); pu, pi = pu[perm], pi[perm]
    indptr, sitems = OP.build_csr(pu, pi, U)
    orc2 = OB.BPROracle(U, I, d, seed=43)
    for step in range(3):
        us, ps, ns = [], [], []
        for r in range(world):
            b = 2 * step + r
            sl = slice(b * B, (b + 1) * B)
            us.append(pu[sl]); ps.append(pi[sl]); ns.append(OP.bpr_negatives(pu[sl], 7, 2, I, indptr, sitems, first_index=b * B))
        orc2.step(np.concatenate(us), np.concatenate(ps), np.concatenate(ns))
    for r in range(world):
        np.testing.assert_allclose(ret[r]["hf_user"], orc2.user, rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(ret[r]["hf_item"], orc2.item, rtol=1e-5, atol=2e-6)
        assert np.isfinite(ret[r]["hf_losses"]).all()
    assert np.array_equal(ret[0]["user"], ret[1]["user"]) and np.array_equal(ret[0]["neumf_W1"], ret[1]["neumf_W1"])
    assert np.array_equal(ret[0]["neumf_uMLP"], ret[1]["neumf_uMLP"])
    Q = (np.random.default_rng(4).integers(-4, 5, size=(77, 64)) / 8.0).astype(np.float32)
    C = (np.random.default_rng(5).integers(-4, 5, size=(1000, 64)) / 8.0).astype(np.float32)
    rv, ri = OT.brute_force_topk(Q, C, 10)
    for r in range(world):
        assert np.array_equal(ret[r]["ti"], ri) and np.array_equal(ret[r]["tv"], rv)
