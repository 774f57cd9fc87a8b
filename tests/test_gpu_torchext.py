"""The two faces of the boundary against each other: every `torch.ops.brk.*` custom op (csrc_torch/brk_torch.cpp, the
PyTorch extension BASELINE.json's north_star names) must give the bits the ctypes binding of the same C-ABI entry point
gives, and must refuse CPU tensors (no fallback)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from binrec_b200 import _torchext as T
    o = T.ops()
    if o is None:
        pytest.fail("libbrk_torch.so is not built / loaded: __graft_entry__.build() builds it")
    return o


def test_ops_match_the_ctypes_binding(dev):
    from binrec_b200 import _native as N, hotpath as H
    o = _ops()
    ctx = N.ctx(dev)
    rng = np.random.default_rng(0)
    rows, d, n = 5000, 64, 3000
    table = torch.from_numpy(rng.standard_normal((rows, d)).astype(np.float32)).to(dev)
    ids = torch.from_numpy((rows * rng.random(n) ** 2).astype(np.int32)).to(dev)
    a = o.gather_rows(table, ids)
    b = torch.empty_like(a)
    N.check(N.lib().brk_gather_rows(ctx, N.ptr(table), rows, d, N.ptr(ids), n, N.ptr(b), N.stream_ptr()), "gather")
    assert torch.equal(a, b) and torch.equal(a, table[ids.long()])
    vals = torch.from_numpy(rng.integers(-4, 5, size=(n, d)).astype(np.float32)).to(dev)
    acc1 = torch.zeros(rows, d, device=dev); acc2 = torch.zeros(rows, d, device=dev)
    t1 = torch.zeros((rows + 31) // 32, dtype=torch.int32, device=dev); t2 = torch.zeros_like(t1)
    o.scatter_add_rows(acc1, ids, vals, t1, 0)
    N.check(N.lib().brk_scatter_add_rows(ctx, N.ptr(acc2), rows, d, N.ptr(ids), n, N.ptr(vals), N.ptr(t2), 0, N.stream_ptr()), "scatter")
    assert torch.equal(acc1, acc2) and torch.equal(t1, t2)
    # scoring + top-K on exact-arithmetic inputs, and the merge of two item-range shards
    Q = torch.from_numpy((rng.integers(-4, 5, size=(300, 64)) / 8.0).astype(np.float32)).to(dev)
    Cm = torch.from_numpy((rng.integers(-4, 5, size=(2000, 64)) / 8.0).astype(np.float32)).to(dev)
    qb, cb = o.rows_to_bf16(Q), o.rows_to_bf16(Cm)
    v, ix = o.score_topk(qb, cb, 10, 0)
    ref = (Q @ Cm.T)
    order = torch.argsort(-ref, dim=1, stable=True)[:, :10]
    assert torch.equal(ix.long(), order) and torch.equal(v, torch.gather(ref, 1, order))
    v0, i0 = o.score_topk(qb, o.rows_to_bf16(Cm[:1000].contiguous()), 10, 0)
    v1, i1 = o.score_topk(qb, o.rows_to_bf16(Cm[1000:].contiguous()), 10, 1000)
    mv, mi = o.topk_merge(torch.stack([v0, v1]), torch.stack([i0, i1]))
    assert torch.equal(mi, ix) and torch.equal(mv, v)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        o.gather_rows(table.cpu(), ids.cpu())


def test_fused_steps_through_the_ops_match_the_oracle(dev):
    """One BPR step (bpr_fwd_bwd + adam_dense_keras ops) and one NeuMF training step (neumf_train_step op, what
    NeuMFNet.train_on_batch calls) against the oracles."""
    from binrec_b200 import hotpath as H
    from binrec_b200.NeuMFModel import NeuMFNet
    from oracle import bpr as OB, neumf as ON
    o = _ops()
    U, I, d, B = 300, 200, 64, 1024
    rng = np.random.default_rng(1)
    orc = OB.BPROracle(U, I, d, seed=42)
    user = H.Table(torch.from_numpy(orc.user.copy()).to(dev)); item = H.Table(torch.from_numpy(orc.item.copy()).to(dev))
    opt = H.Adam(1e-3, device=dev)
    u = rng.integers(0, U, B).astype(np.int32); p = rng.integers(0, I, B).astype(np.int32); n = rng.integers(0, I, B).astype(np.int32)
    loss = o.bpr_fwd_bwd(user.w, user.g, user.touched, item.w, item.g, item.touched,
                         *(torch.from_numpy(x).to(dev) for x in (u, p, n)), 0)
    o.adam_dense_keras([user.w, item.w], [user.g, item.g], [user.m, item.m], [user.v, item.v], 1e-3, 0.9, 0.999, 1e-7, opt.state, True)
    lref = orc.step(u, p, n)
    assert abs(float(loss.item()) - float(lref)) < 1e-6
    np.testing.assert_allclose(user.w.cpu().numpy(), orc.user, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(item.w.cpu().numpy(), orc.item, rtol=1e-5, atol=1e-6)
    assert int(opt.state[0].item()) == 1 and not user.g.any().item()
    net = NeuMFNet(U, I, 16, dropout=0.2, device=dev)               # fp32 tiled instance through the op
    no = ON.NeuMFOracle(U, I, emb=16, dropout=0.2)
    uu = rng.integers(0, U, 512).astype(np.int32); ii = rng.integers(0, I, 512).astype(np.int32)
    yy = (rng.random(512) < 0.3).astype(np.float32)
    lg, _ = net.train_on_batch(*(torch.from_numpy(x).to(dev) for x in (uu, ii, yy)), first_index=7, epoch=2)
    lr, _ = no.step(uu, ii, yy, first_index=7, epoch=2)
    np.testing.assert_allclose(lg.item(), lr, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(net.uMLP.w.cpu().numpy(), no.p.t["uMLP"].detach().numpy(), rtol=1e-5, atol=2e-6)
