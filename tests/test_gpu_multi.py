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
        # fit() under data parallelism: global batch 256 split over the ranks, ragged last batch, two epochs with fresh
        # negatives and a shuffled batch order; then the checkpoint handoff (Adam moments are sharded over the ranks)
        net3 = BPRNet(U, I, d, seed=44, device=dev)
        net3.fit({'customerId_input': pu, 'pProduct_input': pi}, batch_size=256, epochs=2, shuffle=True, sampler_seed=9)
        sd = net3.state_dict()
        net4 = BPRNet(U, I, d, seed=45, device=dev)
        net4.load_state_dict(sd)
        sd4 = net4.state_dict()
        assert all(torch.equal(sd[k], sd4[k]) for k in sd), "state_dict -> load_state_dict -> state_dict is not the identity"
        # NeuMF data parallel through train_steps (the fit loop): He et al. variant has no BatchNorm, so W ranks on slices of
        # the global batch must reproduce the single-process run on the whole batch
        nm2 = NeuMFNet(U, I, 16, dropout=0.0, device=dev, mf_dim=4, mf_mode="hadamard", batch_norm=False)
        rng3 = np.random.default_rng(12)
        nrows = 1000                                           # global batch 192: five full batches + 40 rows (odd split)
        fu = torch.from_numpy(rng3.integers(0, U, nrows).astype(np.int32)).to(dev)
        fi = torch.from_numpy(rng3.integers(0, I, nrows).astype(np.int32)).to(dev)
        fy = torch.from_numpy((rng3.random(nrows) < 0.3).astype(np.float32)).to(dev)
        nm2.train_steps(fu, fi, fy, 192, np.arange(6), epoch=1)
        # two-tower mirrored step: per-replica in-batch softmax (each rank's negatives are its own batch, MirroredStrategy's
        # per-replica loss), SUM-reduced gradients summed over the ranks through the NVLink peer all-reduce, Adagrad
        from binrec_b200.twoTower import TwoTowerModel
        tt = TwoTowerModel(32, I, U, "u", "i", list(range(U)), list(range(I)), semb=16, device=dev)
        tt.compile("Adagrad", learningRate=0.1)
        rng4 = np.random.default_rng(21)
        for step in range(3):
            tu = rng4.integers(2, U + 2, world * 128).astype(np.int32); ti_ = rng4.integers(2, I + 2, world * 128).astype(np.int32)
            lo, hi = D.local_slice(world * 128)
            tt._train_ids(torch.from_numpy(tu[lo:hi]).to(dev), torch.from_numpy(ti_[lo:hi]).to(dev), None)
        if tt._reducer is not None:
            tt._reducer.check()
        torch.cuda.synchronize()
        ret[rank] = dict(user=net.user.w.cpu().numpy(), item=net.item.w.cpu().numpy(), losses=losses,
                         tt_Eu=tt.userTower.emb.w.cpu().numpy(), tt_Wi=tt.itemTower.W.cpu().numpy(), tt_peer=tt._reducer is not None,
                         fit_user=net3.user.w.cpu().numpy(), fit_item=net3.item.w.cpu().numpy(), fit_hist=list(net3.history["loss"]),
                         fit_user_m=sd["user_m"].numpy(), fit_item_v=sd["item_v"].numpy(), fit_t=int(sd["opt_state"][0]),
                         he_W1=nm2.param("W1").cpu().numpy(), he_uMLP=nm2.uMLP.w.cpu().numpy(), he_iMF=nm2.iMF.w.cpu().numpy(),
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
    # fit() through the data-parallel path against the oracle's single-process loop over the same global batches
    orc3 = OB.BPROracle(U, I, d, seed=44)
    n_batches = (len(pu) + 255) // 256
    order_rng = np.random.Generator(np.random.Philox(key=9 + 1000003))
    hist = []
    for e in range(2):
        neg = OP.bpr_negatives(pu, 9, e, I, indptr, sitems)
        ls = []
        for b in order_rng.permutation(n_batches):
            sl = slice(b * 256, min(len(pu), (b + 1) * 256))
            ls.append(orc3.step(pu[sl], pi[sl], neg[sl]))
        hist.append(float(np.mean(ls)))
    for r in range(world):
        np.testing.assert_allclose(ret[r]["fit_user"], orc3.user, rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ret[r]["fit_item"], orc3.item, rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ret[r]["fit_hist"], hist, rtol=1e-5)
        np.testing.assert_allclose(ret[r]["fit_user_m"], orc3.mu, rtol=1e-3, atol=1e-8)
        np.testing.assert_allclose(ret[r]["fit_item_v"], orc3.vi, rtol=1e-3, atol=1e-12)
        assert ret[r]["fit_t"] == 2 * n_batches
    assert np.array_equal(ret[0]["fit_user"], ret[1]["fit_user"])
    from oracle import neumf as ON
    orc4 = ON.NeuMFOracle(U, I, emb=16, dropout=0.0, mf_dim=4, mf_mode="hadamard", batch_norm=False)
    rng3 = np.random.default_rng(12)
    fu = rng3.integers(0, U, 1000).astype(np.int32); fi = rng3.integers(0, I, 1000).astype(np.int32)
    fy = (rng3.random(1000) < 0.3).astype(np.float32)
    for b in range(6):
        sl = slice(b * 192, min(1000, (b + 1) * 192))
        orc4.step(fu[sl], fi[sl], fy[sl], first_index=b * 192, epoch=1)
    ref4 = orc4.p.numpy()
    for r in range(world):
        np.testing.assert_allclose(ret[r]["he_W1"], ref4["W1"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ret[r]["he_uMLP"], ref4["uMLP"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(ret[r]["he_iMF"], ref4["iMF"], rtol=1e-4, atol=2e-5)
    # two-tower: the oracle with the two replicas' losses summed
    from oracle import twotower as OTT
    o5 = OTT.TwoTowerOracle(U, I, 32, 16, seed=42)
    rng4 = np.random.default_rng(21)
    for step in range(3):
        tu = rng4.integers(2, U + 2, world * 128).astype(np.int32); ti_ = rng4.integers(2, I + 2, world * 128).astype(np.int32)
        for v in o5.t.values():
            v.grad = None
        tot = sum(o5.loss(tu[r * 128:(r + 1) * 128], ti_[r * 128:(r + 1) * 128], cand_ids=ti_[r * 128:(r + 1) * 128]) for r in range(world))
        tot.backward()
        with torch.no_grad():
            for k, w in o5.t.items():
                g = w.grad if w.grad is not None else torch.zeros_like(w)
                o5.acc[k] += g * g
                w -= o5.lr * g / (o5.acc[k].sqrt() + 1e-7)
    for r in range(world):
        assert ret[r]["tt_peer"], "the two-tower gradient sum should run over NVLink peer memory on this box"
        np.testing.assert_allclose(ret[r]["tt_Eu"], o5.t["Eu"].detach().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(ret[r]["tt_Wi"], o5.t["Wi"].detach().numpy(), rtol=1e-4, atol=1e-5)
    assert np.array_equal(ret[0]["tt_Eu"], ret[1]["tt_Eu"])
    assert np.array_equal(ret[0]["user"], ret[1]["user"]) and np.array_equal(ret[0]["neumf_W1"], ret[1]["neumf_W1"])
    assert np.array_equal(ret[0]["neumf_uMLP"], ret[1]["neumf_uMLP"])
    Q = (np.random.default_rng(4).integers(-4, 5, size=(77, 64)) / 8.0).astype(np.float32)
    C = (np.random.default_rng(5).integers(-4, 5, size=(1000, 64)) / 8.0).astype(np.float32)
    rv, ri = OT.brute_force_topk(Q, C, 10)
    for r in range(world):
        assert np.array_equal(ret[r]["ti"], ri) and np.array_equal(ret[r]["tv"], rv)
