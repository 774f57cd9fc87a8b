"""GPU parity (through the C ABI) for gather / scatter-add / Philox / fused BPR / optimizers vs the
CPU oracle.  Bit-exact for byte and index work; fp32 tolerance rtol 1e-5 / atol 1e-6 for losses and
updated rows after one step (atomic-order noise), as stated in SURVEY.md section 8c."""
import numpy as np
import pytest
import torch

from oracle import bpr as OB
from oracle import embedding as OE
from oracle import philox as OP

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6


def H():
    from binrec_b200 import hotpath
    return hotpath


@pytest.mark.parametrize("d", [4, 8, 10, 32, 64, 75, 128, 256, 350])
@pytest.mark.parametrize("n", [1, 33, 5000])
def test_gather_rows_bit_exact(dev, d, n):
    rng = np.random.default_rng(d * 1000 + n)
    table = rng.standard_normal((997, d)).astype(np.float32)
    ids = rng.integers(0, 997, n).astype(np.int32)
    out = H().gather_rows(torch.from_numpy(table).to(dev), torch.from_numpy(ids).to(dev))
    assert np.array_equal(out.cpu().numpy(), OE.gather_rows(table, ids))


def test_gather_rows_empty_and_errors(dev):
    t = torch.zeros(8, 16, device=dev)
    out = H().gather_rows(t, torch.zeros(0, dtype=torch.int32, device=dev))
    assert out.shape == (0, 16)
    with pytest.raises(TypeError):
        H().gather_rows(t, torch.zeros(4, dtype=torch.int64, device=dev))


@pytest.mark.parametrize("d", [8, 10, 64, 128])
def test_scatter_add_rows_exact_on_integer_values(dev, d):
    # integer-valued fp32 so that the atomic sum is exact regardless of order => bit-exact
    rng = np.random.default_rng(d)
    rows, n = 211, 20000
    ids = (rows * rng.random(n) ** 3).astype(np.int32)        # heavy head: many duplicates
    vals = rng.integers(-8, 9, size=(n, d)).astype(np.float32)
    acc = torch.zeros(rows, d, device=dev)
    touched = torch.zeros((rows + 31) // 32, dtype=torch.int32, device=dev)
    H().scatter_add_rows(acc, torch.from_numpy(ids).to(dev), torch.from_numpy(vals).to(dev), touched)
    assert np.array_equal(acc.cpu().numpy(), OE.scatter_add_rows(rows, ids, vals))
    bits = np.unpackbits(touched.cpu().numpy().view(np.uint8), bitorder="little")[:rows]
    expect = np.zeros(rows, dtype=np.uint8); expect[np.unique(ids)] = 1
    assert np.array_equal(bits, expect)


@pytest.mark.parametrize("d", [8, 64, 128])
@pytest.mark.parametrize("n", [1, 1000, 20000, 70001])
def test_scatter_add_sorted_mode_exact_on_integer_values(dev, d, n):
    # mode 1 (per-CTA sort + segment-reduce) must give the same sums and touched bits as the definition
    rng = np.random.default_rng(d + n)
    rows = 211
    ids = (rows * rng.random(n) ** 3).astype(np.int32)
    vals = rng.integers(-8, 9, size=(n, d)).astype(np.float32)
    acc = torch.zeros(rows, d, device=dev)
    touched = torch.zeros((rows + 31) // 32, dtype=torch.int32, device=dev)
    H().scatter_add_rows(acc, torch.from_numpy(ids).to(dev), torch.from_numpy(vals).to(dev), touched, mode=1)
    assert np.array_equal(acc.cpu().numpy(), OE.scatter_add_rows(rows, ids, vals))
    bits = np.unpackbits(touched.cpu().numpy().view(np.uint8), bitorder="little")[:rows]
    expect = np.zeros(rows, dtype=np.uint8); expect[np.unique(ids)] = 1
    assert np.array_equal(bits, expect)


def test_scatter_add_auto_mode_picks_by_skew(dev):
    g = torch.Generator(device=dev); g.manual_seed(0)
    n = 100000
    uniform = torch.randint(0, 6040, (n,), generator=g, device=dev, dtype=torch.int32)
    skewed = (6040 * torch.rand(n, generator=g, device=dev) ** 3).to(torch.int32)
    assert H().index_skew(uniform) < 0.01 <= H().index_skew(skewed)
    vals = torch.ones(n, 64, device=dev)
    for ids in (uniform, skewed):
        acc = torch.zeros(6040, 64, device=dev)
        H().scatter_add_rows(acc, ids, vals, mode="auto")
        assert torch.equal(acc[:, 0], torch.bincount(ids.long(), minlength=6040).float())


def test_scatter_add_is_gather_transpose_property(dev):
    # <gather(T, ids), V> == <T, scatter_add(ids, V)>  (linearity / adjointness), exact in integers
    rng = np.random.default_rng(5)
    rows, d, n = 64, 32, 4096
    T = rng.integers(-4, 5, size=(rows, d)).astype(np.float32)
    V = rng.integers(-4, 5, size=(n, d)).astype(np.float32)
    ids = rng.integers(0, rows, n).astype(np.int32)
    Td, Vd, idd = (torch.from_numpy(x).to(dev) for x in (T, V, ids))
    lhs = (H().gather_rows(Td, idd).double() * Vd.double()).sum().item()
    acc = torch.zeros(rows, d, device=dev)
    H().scatter_add_rows(acc, idd, Vd)
    rhs = (Td.double() * acc.double()).sum().item()
    assert lhs == rhs


def test_philox_known_answers_on_device(dev):
    ctr = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]],
                   dtype=np.uint32)
    exp = [[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8], [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd],
           [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]]
    keys = [(0, 0), (0xffffffff, 0xffffffff), (0xa4093822, 0x299f31d0)]
    for i in range(3):
        c = torch.from_numpy(ctr[i:i + 1].view(np.int32)).to(dev)
        out = H().philox4x32_10(c, keys[i][0], keys[i][1]).cpu().numpy().view(np.uint32)[0]
        assert [int(x) for x in out] == exp[i]


def test_philox_raw_matches_oracle_bulk(dev):
    rng = np.random.default_rng(9)
    ctr = rng.integers(0, 2**32, size=(10000, 4), dtype=np.uint64).astype(np.uint32)
    out = H().philox4x32_10(torch.from_numpy(ctr.view(np.int32)).to(dev), 20261018, 7)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), OP.philox4x32_10(ctr, (20261018, 7)))


@pytest.mark.parametrize("U,I,P", [(50, 40, 600), (300, 17, 3000), (6040, 3706, 200000)])
def test_philox_bpr_negatives_bit_exact(dev, U, I, P):
    rng = np.random.default_rng(U)
    key = np.unique(rng.integers(0, U * I, P))
    pu, pi = (key // I).astype(np.int32), (key % I).astype(np.int32)
    indptr, sitems = OP.build_csr(pu, pi, U)
    for epoch, first in ((0, 0), (3, 12345)):
        ref = OP.bpr_negatives(pu, 7, epoch, I, indptr, sitems, first_index=first)
        got = H().philox_bpr_negatives(torch.from_numpy(pu).to(dev), 7, epoch, I,
                                       torch.from_numpy(indptr).to(dev), torch.from_numpy(sitems).to(dev),
                                       first_index=first)
        assert np.array_equal(got.cpu().numpy(), ref)


def test_philox_neumf_negatives_bit_exact(dev):
    rng = np.random.default_rng(2)
    pu = rng.integers(0, 6040, 100000).astype(np.int32); pi = rng.integers(0, 3706, 100000).astype(np.int32)
    ru, ri = OP.neumf_negatives(pu, pi, 400000, 7, 2, first_index=99)
    gu, gi = H().philox_neumf_negatives(torch.from_numpy(pu).to(dev), torch.from_numpy(pi).to(dev), 400000, 7, 2,
                                        first_index=99)
    assert np.array_equal(gu.cpu().numpy(), ru) and np.array_equal(gi.cpu().numpy(), ri)


def _bpr_setup(dev, U, I, d, B, seed=0, skew=True):
    rng = np.random.default_rng(seed)
    orc = OB.BPROracle(U, I, d, seed=42)
    # larger weights than the Keras init so that sigmoid is exercised away from 0.5
    orc.user = (orc.user * 8).astype(np.float32); orc.item = (orc.item * 8).astype(np.float32)
    u = (U * rng.random(B) ** (2.0 if skew else 1.0)).astype(np.int32)
    p = (I * rng.random(B) ** (2.0 if skew else 1.0)).astype(np.int32)
    n = rng.integers(0, I, B).astype(np.int32)
    h = H()
    user = h.Table(torch.from_numpy(orc.user.copy()).to(dev))
    item = h.Table(torch.from_numpy(orc.item.copy()).to(dev))
    return orc, user, item, u, p, n


@pytest.mark.parametrize("d", [8, 32, 64, 128, 256, 10, 350])
def test_bpr_fwd_bwd_one_step_keras_adam(dev, d):
    orc, user, item, u, p, n = _bpr_setup(dev, 300, 200, d, 2048)
    h = H()
    opt = h.Adam(1e-3, device=dev)
    ud, pd, nd = (torch.from_numpy(x).to(dev) for x in (u, p, n))
    loss_ref, gu, gi = OB.bpr_loss_and_grads(orc.user, orc.item, u, p, n)
    loss = h.bpr_fwd_bwd(user, item, ud, pd, nd)
    np.testing.assert_allclose(loss.item(), loss_ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(user.g.cpu().numpy(), gu, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(item.g.cpu().numpy(), gi, rtol=1e-4, atol=1e-7)
    x = h.bpr_scores(user.w, item.w, ud, pd, nd)
    np.testing.assert_allclose(x.cpu().numpy(), OB.bpr_forward(orc.user, orc.item, u, p, n)[0], rtol=1e-5, atol=1e-6)
    opt.apply([user, item])
    orc.step(u, p, n)
    np.testing.assert_allclose(user.w.cpu().numpy(), orc.user, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(item.w.cpu().numpy(), orc.item, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(user.m.cpu().numpy(), orc.mu, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(user.v.cpu().numpy(), orc.vu, rtol=1e-4, atol=1e-12)
    assert not user.g.any().item() and not item.g.any().item()       # accumulators re-zeroed
    assert not user.touched.any().item() and not item.touched.any().item()
    assert opt.step.item() == 1


@pytest.mark.parametrize("sparse", ["keras", "lazy"])
def test_bpr_ten_steps_match_oracle(dev, sparse):
    U, I, d, B = 500, 400, 64, 1024
    orc, user, item, *_ = _bpr_setup(dev, U, I, d, B)
    orc.optimizer = "adam_keras" if sparse == "keras" else "adam_lazy"
    h = H()
    opt = h.Adam(1e-3, sparse=sparse, device=dev)
    rng = np.random.default_rng(11)
    for step in range(10):
        u = rng.integers(0, U, B).astype(np.int32); p = rng.integers(0, I, B).astype(np.int32)
        n = rng.integers(0, I, B).astype(np.int32)
        loss = h.bpr_fwd_bwd(user, item, *(torch.from_numpy(x).to(dev) for x in (u, p, n)))
        opt.apply([user, item])
        lref = orc.step(u, p, n)
        np.testing.assert_allclose(loss.item(), lref, rtol=RTOL, atol=ATOL)
    assert opt.step.item() == 10
    np.testing.assert_allclose(user.w.cpu().numpy(), orc.user, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(item.w.cpu().numpy(), orc.item, rtol=1e-4, atol=2e-6)
    assert not user.touched.any().item() and not item.g.any().item()


def test_adagrad_rows_and_dense_agree_with_oracle(dev):
    h = H()
    rng = np.random.default_rng(4)
    rows, d, n = 3000, 128, 700
    w0 = rng.standard_normal((rows, d)).astype(np.float32)
    ids = rng.integers(0, rows, n).astype(np.int32)
    vals = rng.standard_normal((n, d)).astype(np.float32)
    ref_w = w0.copy(); ref_acc = np.full_like(w0, 0.1)
    OE.adagrad_rows(ref_w, ref_acc, OE.scatter_add_rows(rows, ids, vals), np.unique(ids), lr=0.1)
    for thr in (0, 1 << 40):      # force the row-sparse kernel, then the dense kernel
        t = h.Table(torch.from_numpy(w0.copy()).to(dev), slots=1, slot_init=0.1)
        h.scatter_add_rows(t.g, torch.from_numpy(ids).to(dev), torch.from_numpy(vals).to(dev), t.touched)
        h.Adagrad(0.1, rows_threshold_bytes=thr).apply([t])
        np.testing.assert_allclose(t.w.cpu().numpy(), ref_w, rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(t.m.cpu().numpy(), ref_acc, rtol=RTOL, atol=ATOL)
        assert not t.g.any().item()


def test_bpr_multi_step_cooperative_kernel_matches_oracle(dev):
    """brk_bpr_train_steps: K steps inside one cooperative kernel (fwd/bwd -> grid.sync -> Adam) must
    equal K oracle steps, including a ragged last batch and a permuted batch order."""
    from binrec_b200.BPRModel import BPRNet
    U, I, d, B = 400, 300, 64, 512
    rng = np.random.default_rng(21)
    total = 5 * B + 77                                  # ragged last batch
    u = rng.integers(0, U, total).astype(np.int32); p = rng.integers(0, I, total).astype(np.int32)
    n = rng.integers(0, I, total).astype(np.int32)
    net = BPRNet(U, I, d, seed=42, device=dev)
    orc = OB.BPROracle(U, I, d, seed=42)
    net.set_training_pairs(u, p); net.set_negatives(n)
    order = [3, 0, 5, 1, 4, 2, 5, 0]
    losses = net.train_steps(order, B)
    ref = []
    for b in order:
        sl = slice(b * B, min(total, (b + 1) * B))
        ref.append(orc.step(u[sl], p[sl], n[sl]))
    np.testing.assert_allclose(losses.cpu().numpy(), np.array(ref), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(net.user.w.cpu().numpy(), orc.user, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(net.item.w.cpu().numpy(), orc.item, rtol=1e-4, atol=2e-6)
    assert net.optimizer.step.item() == len(order) and not net.grad_arena.any().item()
    # the separate-kernel path (BRK_NO_COOP) gives the same result
    import os
    net2 = BPRNet(U, I, d, seed=42, device=dev); net2.set_training_pairs(u, p); net2.set_negatives(n)
    os.environ["BRK_NO_COOP"] = "1"
    try:
        l2 = net2.train_steps(order, B)
    finally:
        del os.environ["BRK_NO_COOP"]
    np.testing.assert_allclose(l2.cpu().numpy(), losses.cpu().numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(net2.user.w.cpu().numpy(), net.user.w.cpu().numpy(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("path", ["copy", "copy_packed", "mapped"])
def test_bpr_host_fed_steps_match_oracle(dev, path):
    """BPRNet.train_steps_from_host (pinned host ids -> H2D on a copy stream through a ring of staging
    slots -> cooperative step kernel drawing its own Philox negatives -> loss D2H) and the zero-copy
    train_steps_mapped (one launch, ids pulled over PCIe by the kernel) against the oracle with the
    oracle's sampler."""
    from binrec_b200.BPRModel import BPRNet
    U, I, d, B = 300, 200, 64, 256
    rng = np.random.default_rng(31)
    key = np.unique(rng.integers(0, U * I, 4000))
    u, p = (key // I).astype(np.int32), (key % I).astype(np.int32)
    perm = rng.permutation(len(u)); u, p = u[perm], p[perm]
    net = BPRNet(U, I, d, seed=42, device=dev)
    net.set_training_pairs(u, p)
    orc = OB.BPROracle(U, I, d, seed=42)
    indptr, sitems = OP.build_csr(u, p, U)
    # ragged tail batches; 37 steps = chunks of 2, 4, 8, 16 and 7 steps of the host-fed ring (ramped plan of brk_bpr_train_steps_host)
    order = [2, 0, 7, 3, 3, 1, len(u) // B, 5, 4, 6, 0, 1, 2, len(u) // B, 9, 8, 3, 7, 10] + [(3 * k + 1) % (len(u) // B + 1) for k in range(18)]
    hu, hp = torch.from_numpy(u).pin_memory(), torch.from_numpy(p).pin_memory()
    if path == "copy":
        losses = net.train_steps_from_host(hu, hp, order, B, 7, 5)
    elif path == "copy_packed":
        losses = net.train_steps_from_host(BPRNet.pack_host_batches(u, p, B), None, order, B, 7, 5)
    else:
        losses = net.train_steps_mapped(hu, hp, order, B, 7, 5)
    torch.cuda.synchronize()
    ref = []
    for b in order:
        sl = slice(b * B, min(len(u), (b + 1) * B))
        neg = OP.bpr_negatives(u[sl], 7, 5, I, indptr, sitems, first_index=b * B)
        ref.append(orc.step(u[sl], p[sl], neg))
    np.testing.assert_allclose(losses.numpy(), np.array(ref), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(net.user.w.cpu().numpy(), orc.user, rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(net.item.w.cpu().numpy(), orc.item, rtol=1e-4, atol=2e-6)
