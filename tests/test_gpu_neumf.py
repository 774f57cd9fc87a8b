"""GPU parity of the fused NeuMF forward/backward (through the C ABI) against the torch-autograd
oracle (oracle/neumf.py) on identical inputs and initial weights.

Stated fp32 tolerances: loss/predictions rtol 1e-5 / atol 1e-6; gradients rtol 1e-3 / atol 1e-6
(reduction-order noise of the batch sums, the BN-backward cancellation and the atomic accumulation); weights after one Keras-Adam
step rtol 1e-5 / atol 2e-6, after 5 steps atol 1e-5."""
import numpy as np
import pytest
import torch

from oracle import neumf as ON
from oracle import topk as OT

pytestmark = pytest.mark.gpu


def _mk(dev, E, hidden, act, loss, dropout, U=300, I=200, seed=42, **variant):
    from binrec_b200.NeuMFModel import NeuMFNet
    orc = ON.NeuMFOracle(U, I, emb=E, hidden=hidden, seed=seed, act=act, loss=loss, dropout=dropout, dropout_seed=11, **variant)
    net = NeuMFNet(U, I, E, hidden=hidden, act=act, loss=loss, dropout=dropout, seed=seed, dropout_seed=11, device=dev, **variant)
    # identical initial weights (same draw order) -- verify instead of assuming
    ref = orc.p.numpy()
    for name, tab in zip(("uMLP", "iMLP", "uMF", "iMF"), net.tables()):
        assert np.array_equal(tab.w.cpu().numpy(), ref[name])
    for name in net.DENSE_ORDER:
        assert np.array_equal(net.param(name).cpu().numpy().reshape(ref[name].shape), ref[name])
    return orc, net


def _batch(rng, U, I, B):
    u = (U * rng.random(B) ** 2).astype(np.int32)
    i = (I * rng.random(B) ** 2).astype(np.int32)
    y = (rng.random(B) < 0.25).astype(np.float32)
    return u, i, y


SPECS = [(8, (8, 4, 2), "relu", "mse"), (32, (32, 16, 8), "relu", "mse"), (64, (64, 32, 16), "relu", "mse"),
         (16, (16, 8, 4), "sigmoid", "bce"), (10, (100, 50, 10), "sigmoid", "bce"),
         # numFactor is a free attribute of the reference model (RModel.py:35): widths without a tiled instance run on
         # the any-width kernels (csrc/neumf_generic.cu)
         (20, None, "relu", "mse"), (7, None, "relu", "mse"), (48, None, "sigmoid", "bce"), (128, None, "relu", "mse")]


@pytest.mark.parametrize("E,hidden,act,loss", SPECS)
@pytest.mark.parametrize("dropout", [0.0, 0.2])
@pytest.mark.parametrize("B", [1000, 128])
def test_neumf_forward_backward_matches_autograd(dev, E, hidden, act, loss, dropout, B):
    U, I = 300, 200
    orc, net = _mk(dev, E, hidden, act, loss, dropout, U, I)
    rng = np.random.default_rng(B + E)
    u, i, y = _batch(rng, U, I, B)
    first = 4096
    lref, oref, aux = orc.loss_and_grads(u, i, y, first_index=first, epoch=3)
    ud, idd, yd = (torch.from_numpy(x).to(dev) for x in (u, i, y))
    lgot, ogot = net.forward_backward(ud, idd, yd, first_index=first, epoch=3)
    np.testing.assert_allclose(ogot.cpu().numpy(), oref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lgot.item(), float(lref), rtol=1e-5, atol=1e-6)
    for name, tab in zip(("uMLP", "iMLP", "uMF", "iMF"), net.tables()):
        np.testing.assert_allclose(tab.g.cpu().numpy(), orc.p.t[name].grad.numpy(), rtol=1e-3, atol=1e-6, err_msg=name)
    for name in net.DENSE_ORDER:
        g = orc.p.t[name].grad.numpy()
        np.testing.assert_allclose(net.param(name, grad=True).cpu().numpy().reshape(g.shape), g, rtol=1e-3, atol=1e-6,
                                   err_msg=name)
    # workspace accumulators are left zero for the next step
    assert not net._bufs["acc"].any().item()


@pytest.mark.parametrize("E,mf_dim,act,loss,dropout", [(16, 4, "relu", "mse", 0.0), (20, 8, "sigmoid", "bce", 0.2), (12, 12, "relu", "mse", 0.0)])
def test_neumf_he_variant_any_width_matches_autograd(dev, E, mf_dim, act, loss, dropout):
    """The He et al. variant (Hadamard GMF vector into the head, no BatchNorm) at widths without a tensor-core instance:
    fp32 on the any-width kernels, fp32 tolerance against the autograd oracle."""
    U, I, B = 300, 200, 700
    orc, net = _mk(dev, E, None, act, loss, dropout, U, I, mf_dim=mf_dim, mf_mode="hadamard", batch_norm=False)
    assert not net.tensor_cores
    rng = np.random.default_rng(E)
    u, i, y = _batch(rng, U, I, B)
    lref, oref, _ = orc.loss_and_grads(u, i, y, first_index=64, epoch=1)
    lgot, ogot = net.forward_backward(*(torch.from_numpy(x).to(dev) for x in (u, i, y)), first_index=64, epoch=1)
    np.testing.assert_allclose(ogot.cpu().numpy(), oref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lgot.item(), float(lref), rtol=1e-5, atol=1e-6)
    for name, tab in zip(("uMLP", "iMLP", "uMF", "iMF"), net.tables()):
        np.testing.assert_allclose(tab.g.cpu().numpy(), orc.p.t[name].grad.numpy(), rtol=1e-3, atol=1e-6, err_msg=name)
    for name in ("W1", "b1", "W2", "b2", "W3", "b3", "W4", "b4"):
        g = orc.p.t[name].grad.numpy()
        np.testing.assert_allclose(net.param(name, grad=True).cpu().numpy().reshape(g.shape), g, rtol=1e-3, atol=1e-6, err_msg=name)
    for name in ("g1", "be1", "g2", "be2"):                 # unused BatchNorm slots stay untouched
        assert not net.param(name, grad=True).any().item()
    for step in range(3):                                     # and it trains: three Keras-Adam steps against the oracle
        u, i, y = _batch(rng, U, I, B)
        lref, _ = orc.step(u, i, y, first_index=step * B, epoch=0)
        lgot, _ = net.train_on_batch(*(torch.from_numpy(x).to(dev) for x in (u, i, y)), first_index=step * B, epoch=0)
        # Adam normalises every step to ~lr, so rounding noise in gradients near |g| ~ eps moves weights by up to lr
        # per step in both implementations (see test_neumf_five_steps_match_oracle_keras_adam): 5e-4 on the loss
        np.testing.assert_allclose(lgot.item(), lref, rtol=5e-4, atol=1e-6)


@pytest.mark.parametrize("E,hidden,act,loss", [SPECS[1], SPECS[4]])
def test_neumf_five_steps_match_oracle_keras_adam(dev, E, hidden, act, loss):
    """Five Keras-Adam steps.  Adam divides by sqrt(v)+eps, which amplifies sub-1e-7 gradient noise
    wherever |g| is of the order of eps (1e-7), so weights are judged against an fp64 run of the
    oracle: the device must be as close to it as the fp32 oracle is (factor 5 + 2e-6 slack), and
    within atol 2e-4 of the fp32 oracle outright (parameters whose true gradient is ~eps, e.g. a bias
    feeding a BatchNorm, move by noise-dominated steps of up to lr = 1e-3 each in BOTH implementations)."""
    from binrec_b200.NeuMFModel import NeuMFNet
    U, I, B = 300, 200, 512
    orc, net = _mk(dev, E, hidden, act, loss, 0.2, U, I)
    o64 = ON.NeuMFOracle(U, I, emb=E, hidden=hidden, act=act, loss=loss, dropout=0.2, dropout_seed=11,
                         dtype=torch.float64)
    rng = np.random.default_rng(5)
    for step in range(5):
        u, i, y = _batch(rng, U, I, B)
        lref, _ = orc.step(u, i, y, first_index=step * B, epoch=0)
        o64.step(u, i, y, first_index=step * B, epoch=0)
        lgot, _ = net.train_on_batch(*(torch.from_numpy(x).to(dev) for x in (u, i, y)), first_index=step * B, epoch=0)
        np.testing.assert_allclose(lgot.item(), lref, rtol=1e-4, atol=1e-6)
    ref, ref64 = orc.p.numpy(), o64.p.numpy()
    got = {name: tab.w.cpu().numpy() for name, tab in zip(("uMLP", "iMLP", "uMF", "iMF"), net.tables())}
    got.update({name: net.param(name).cpu().numpy().reshape(ref[name].shape) for name in net.DENSE_ORDER})
    for name, w in got.items():
        err_dev = np.abs(w - ref64[name]).max()
        err_o32 = np.abs(ref[name] - ref64[name]).max()
        assert err_dev <= 5 * err_o32 + 2e-5, (name, err_dev, err_o32)
        np.testing.assert_allclose(w, ref[name], rtol=1e-4, atol=2e-4, err_msg=name)
    h1, h2, _ = hidden
    bn = net.bn_moving.cpu().numpy()
    for g, want in ((bn[:h1], orc.p.mm1), (bn[h1:2 * h1], orc.p.mv1), (bn[2 * h1:2 * h1 + h2], orc.p.mm2),
                    (bn[2 * h1 + h2:], orc.p.mv2)):
        # moving statistics follow weights that went through five noise-amplifying Adam steps (see the docstring):
        # the multi-step fp32 tolerance of SURVEY.md section 8c (1e-4), not the one-step 1e-5
        np.testing.assert_allclose(g, want.numpy(), rtol=1e-4, atol=1e-6)
    assert net.optimizer.step.item() == 5
    # inference path uses the moving statistics
    u, i, y = _batch(rng, U, I, 777)
    out, l = net.predict_on_batch(*(torch.from_numpy(x).to(dev) for x in (u, i, y)))
    np.testing.assert_allclose(out.cpu().numpy(), orc.predict(u, i), rtol=1e-4, atol=1e-5)
    out2, none = net.predict_on_batch(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev))
    assert none is None and torch.equal(out, out2)


def test_topk_rows_matches_oracle_with_ties(dev):
    from binrec_b200 import hotpath as H
    rng = np.random.default_rng(0)
    for R, I, k in ((7, 5, 10), (33, 3706, 10), (4, 100000, 32), (100, 64, 5)):
        S = (rng.integers(-20, 21, size=(R, I)) / 8.0).astype(np.float32)
        v, ix = H.topk_rows(torch.from_numpy(S).to(dev), k)
        rv, ri = OT.topk_from_scores(S, k)
        assert np.array_equal(ix.cpu().numpy(), ri) and np.array_equal(v.cpu().numpy(), rv)


def test_neumf_model_class_end_to_end(dev, tmp_path):
    from binrec_b200.NeuMFModel import NeuMFModel
    from binrec_b200 import synth
    users, items = synth.make_interactions(num_users=400, num_items=300, num_pos=20000, seed=5)
    m = NeuMFModel(workDir=str(tmp_path))
    m.epochs = 2
    res = m.train((users, items), None)
    assert res["result"] == "completed" and len(res["metrics"]) == 4
    assert m.model.history["loss"][1] < m.model.history["loss"][0]          # it learns
    ds = m.bootstrapDataset((users[:1000], items[:1000]), shuffle=False)
    feats, label = next(iter(ds))
    assert set(feats) == {"user", "item"} and label.shape == feats["user"].shape
    assert abs(float(ds.y.mean().item()) - 0.25) < 1e-6                     # negRatio 3 -> 1 positive in 4
    pds = m.getPredictDataSet(m.getPredictableUsers()[0])               # RModel.py:168-170: the frame of predictForUser
    assert pds.n == 4 * len(m._testProducts) and not pds.shuffle          # bootstrapDataset adds its 3x "negatives" here too
    recs = m.predictForUser(m.getPredictableUsers()[0], 5)
    assert len(recs) == 5 and all(isinstance(a, str) and isinstance(b, str) for a, b in recs)
    scores = [float(b) for _, b in recs]
    assert scores == sorted(scores, reverse=True)
    m2 = NeuMFModel(workDir=str(tmp_path)); m2.compileModel(None, m.model.numUser, m.model.numItem, m.numFactor)
    m2.restoreFromLatestCheckPoint()
    assert torch.equal(m2.model.uMLP.w, m.model.uMLP.w)


@pytest.mark.parametrize("sparse", ["keras", "lazy"])
def test_train_steps_in_one_call_equal_the_per_step_loop(dev, sparse):
    """brk_neumf_train_steps (fit's inner loop as one C call) against train_on_batch per batch: same batches, same
    dropout stream positions, same optimizer calls -- equal up to the RED accumulation order."""
    from binrec_b200.NeuMFModel import NeuMFNet
    rng = np.random.default_rng(3)
    U, I, n, B = 500, 300, 5000, 384                         # ragged last batch: 5000 = 13 * 384 + 8
    u = torch.from_numpy(rng.integers(0, U, n).astype(np.int32)).to(dev)
    i = torch.from_numpy(rng.integers(0, I, n).astype(np.int32)).to(dev)
    y = torch.from_numpy((rng.random(n) < 0.25).astype(np.float32)).to(dev)
    order = rng.permutation((n + B - 1) // B)
    a = NeuMFNet(U, I, 32, dropout=0.2, device=dev, sparse_adam=sparse)
    b = NeuMFNet(U, I, 32, dropout=0.2, device=dev, sparse_adam=sparse)
    la = torch.empty(len(order), device=dev)
    for k, bb in enumerate(order):
        s = slice(int(bb) * B, min(n, (int(bb) + 1) * B))
        a.train_on_batch(u[s], i[s], y[s], first_index=int(bb) * B, epoch=2, loss_out=la[k:k + 1])
    lb = b.train_steps(u, i, y, B, order, epoch=2)
    np.testing.assert_allclose(lb.cpu().numpy(), la.cpu().numpy(), rtol=1e-4, atol=1e-6)
    for ta, tb in zip(a.tables() + [a.dense], b.tables() + [b.dense]):
        np.testing.assert_allclose(tb.w.cpu().numpy(), ta.w.cpu().numpy(), rtol=1e-3, atol=2e-5)
    assert int(a.optimizer.state[0].item()) == int(b.optimizer.state[0].item()) == len(order)
    np.testing.assert_allclose(b.bn_moving.cpu().numpy(), a.bn_moving.cpu().numpy(), rtol=1e-4, atol=1e-6)


def test_host_fed_steps_equal_device_fed_steps(dev):
    """brk_neumf_train_steps_host (frame in pinned host memory, one H2D copy per step on the copy stream, four staging
    slots) against the same batches fed from device tensors: identical losses and weights -- the copies run ahead of the
    compute but every step must see exactly its batch."""
    from binrec_b200.NeuMFModel import NeuMFNet
    rng = np.random.default_rng(4)
    U, I, B, nb = 500, 300, 256, 11
    u = rng.integers(0, U, nb * B).astype(np.int32); i = rng.integers(0, I, nb * B).astype(np.int32)
    y = (rng.random(nb * B) < 0.25).astype(np.float32)
    order = rng.permutation(nb)
    for kw in (dict(tensor_cores=True), dict(), dict(mf_dim=8, mf_mode="hadamard", batch_norm=False, dropout=0.0)):
        kw.setdefault("dropout", 0.2)
        a = NeuMFNet(U, I, 32, device=dev, **kw); b = NeuMFNet(U, I, 32, device=dev, **kw)
        packed = NeuMFNet.pack_host_batches(u, i, y, B)
        la = a.train_steps_from_host(packed, order, epoch=3)
        ud, idd, yd = (torch.from_numpy(x).to(dev) for x in (u, i, y))
        lb = b.train_steps(ud, idd, yd, B, order, epoch=3)
        torch.cuda.synchronize()
        np.testing.assert_allclose(la.numpy(), lb.cpu().numpy(), rtol=1e-5, atol=1e-6)
        for ta, tb in zip(a.tables() + [a.dense], b.tables() + [b.dense]):
            np.testing.assert_allclose(ta.w.cpu().numpy(), tb.w.cpu().numpy(), rtol=1e-4, atol=2e-5)
        assert int(a.optimizer.state[0].item()) == nb


def test_one_launch_step_dense_gradients_are_bit_reproducible(dev):
    """The one-launch step sums BatchNorm statistics and dense gradients through per-tile slots in a fixed order (no
    atomics on shared addresses): two runs on the same batch give bit-identical predictions, loss and dense gradients."""
    from binrec_b200.NeuMFModel import NeuMFNet
    rng = np.random.default_rng(8)
    U, I, B = 6040, 3706, 16384
    u = torch.from_numpy((U * rng.random(B) ** 2).astype(np.int32)).to(dev)
    i = torch.from_numpy((I * rng.random(B) ** 2).astype(np.int32)).to(dev)
    y = torch.from_numpy((rng.random(B) < 0.2).astype(np.float32)).to(dev)
    net = NeuMFNet(U, I, 32, dropout=0.2, device=dev, tensor_cores=True)
    l0, o0 = net.forward_backward(u, i, y, first_index=5, epoch=1)
    g0, l0, o0 = net.dense.g.clone(), l0.clone(), o0.clone()
    net.grad_arena.zero_()
    l1, o1 = net.forward_backward(u, i, y, first_index=5, epoch=1)
    assert torch.equal(o0, o1) and torch.equal(l0, l1) and torch.equal(g0, net.dense.g)
    assert bool(g0.abs().sum() > 0)
