"""Checkpoint / serve handoff on the device: train -> checkpoint directory -> a fresh
process-like object restores and serves; row-sharded runs restore under a different number of shards bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_neumf_restore_on_a_fresh_object_and_batched_serving(dev, tmp_path):
    from binrec_b200.NeuMFModel import NeuMFModel
    from binrec_b200 import synth
    users, items = synth.make_interactions(num_users=400, num_items=300, num_pos=20000, seed=5)
    m = NeuMFModel(workDir=str(tmp_path)); m.epochs = 1
    m.train((users, items), None)
    served = NeuMFModel(workDir=str(tmp_path))                    # what RecommendationEndpoint.py:47-50 does: no compile
    served.restoreFromLatestCheckPoint()
    assert served.getPredictableUsers() == m.getPredictableUsers()
    for name in ("uMLP", "iMLP", "uMF", "iMF", "dense"):
        a, b = getattr(served.model, name), getattr(m.model, name)
        assert torch.equal(a.w, b.w) and torch.equal(a.m, b.m) and torch.equal(a.v, b.v)
    assert torch.equal(served.model.bn_moving, m.model.bn_moving)
    assert torch.equal(served.model.optimizer.state, m.model.optimizer.state)
    some = m.getPredictableUsers()[:17]
    batch = served.predictForUsers(some, 7)
    assert len(batch) == 17 and all(len(r) == 7 for r in batch)
    for u, want in zip(some[:5], batch[:5]):
        assert m.predictForUser(u, 7) == want                     # same model, one user at a time == batched
    assert served.predictForUsers([], 3) == []
    with pytest.raises(ValueError):
        served.predictForUsers([10 ** 6], 3)
    # training resumes from the restored optimizer state: one more identical step on both gives the same weights
    u = torch.from_numpy(users[:512]).to(dev); i = torch.from_numpy(items[:512]).to(dev); y = torch.ones(512, device=dev)
    for mm in (m, served):
        mm.model.train_on_batch(u, i, y, first_index=0, epoch=9)
    np.testing.assert_allclose(served.model.uMF.w.cpu().numpy(), m.model.uMF.w.cpu().numpy(), rtol=1e-5, atol=1e-7)


def test_bpr_restore_and_batched_top_products(dev, tmp_path):
    from binrec_b200.BPRModel import BPRModel, bpr_predict
    from binrec_b200 import synth
    users, items = synth.make_interactions(num_users=300, num_items=500, num_pos=15000, seed=9)
    m = BPRModel(workDir=str(tmp_path)); m.epochs, m.numFactor = 2, 16
    m.train((users, items), None)
    m.saveCheckPoint()
    served = BPRModel(workDir=str(tmp_path))
    served.restoreFromLatestCheckPoint()
    assert torch.equal(served.model.user.w, m.model.user.w) and torch.equal(served.model.item.v, m.model.item.v)
    assert served.productIds == m.productIds
    got = served.predictForUsers([0, 5, 299], 10)
    assert len(got) == 3 and all(len(r) == 10 for r in got)
    for uid, recs in zip([0, 5, 299], got):
        exact = bpr_predict(m.model, uid, np.arange(m.model.item.rows)).cpu().numpy()        # fp32 scores (bpr.py:122-133)
        ids = [int(a) for a, _ in recs]
        vals = np.array([float(b) for _, b in recs])
        assert (np.diff(vals) <= 0).all() and len(set(ids)) == 10
        np.testing.assert_allclose(vals, exact[ids], rtol=2e-2, atol=2e-3)                     # bf16 operands
        kth = np.sort(exact)[-10]
        assert (exact[ids] >= kth - 2e-2 * max(1.0, abs(kth))).all()                          # gap-aware membership


@pytest.mark.parametrize("G_save,G_load", [(2, 3), (3, 1), (1, 2)])
def test_sharded_bpr_checkpoint_restores_under_another_world_size(dev, tmp_path, G_save, G_load):
    from binrec_b200.sharded import ShardedBPRNet
    U, I, d, B = 301, 203, 16, 400
    rng = np.random.default_rng(3)
    batches = [tuple(torch.from_numpy(rng.integers(0, n, B).astype(np.int32)).to(dev) for n in (U, I, I)) for _ in range(4)]
    a = ShardedBPRNet(U, I, d, device=dev, emulate=G_save)
    for u, p, n in batches[:3]:
        a.train_on_batch(u, p, n)
    a.save_checkpoint(str(tmp_path / "cp"))
    b = ShardedBPRNet(U, I, d, device=dev, emulate=G_load, full_init={"user": np.zeros((U, d), np.float32), "item": np.zeros((I, d), np.float32)})
    b.load_checkpoint(str(tmp_path / "cp"))
    assert np.array_equal(a.user.full_weights(), b.user.full_weights())
    assert np.array_equal(a.item.full_weights(), b.item.full_weights())
    assert torch.equal(a.optimizer.state, b.optimizer.state)
    la, lb = a.train_on_batch(*batches[3]), b.train_on_batch(*batches[3])                      # moments came along too
    np.testing.assert_allclose(la.item(), lb.item(), rtol=1e-6)
    np.testing.assert_allclose(a.user.full_weights(), b.user.full_weights(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(a.item.full_weights(), b.item.full_weights(), rtol=1e-5, atol=1e-7)


def test_sharded_neumf_checkpoint_serves_from_an_unsharded_model(dev, tmp_path):
    """The 8-GPU-train / 1-GPU-serve handoff, emulated with 3 shards in one process: the checkpoint written by the
    sharded model loads into NeuMFNet (whole tables) and gives the same predictions."""
    from binrec_b200 import checkpoint as CK
    from binrec_b200.NeuMFModel import NeuMFNet
    from binrec_b200.sharded import ShardedNeuMFNet
    U, I, E, B = 301, 203, 8, 500
    rng = np.random.default_rng(1)
    sh = ShardedNeuMFNet(U, I, E, dropout=0.0, device=dev, mode="peer", emulate=3)
    for _ in range(3):
        u = torch.from_numpy(rng.integers(0, U, B).astype(np.int32)).to(dev)
        i = torch.from_numpy(rng.integers(0, I, B).astype(np.int32)).to(dev)
        y = torch.from_numpy((rng.random(B) < 0.3).astype(np.float32)).to(dev)
        sh.train_on_batch(u, i, y)
    sh.save_checkpoint(str(tmp_path / "cp"))
    net = NeuMFNet(U, I, E, dropout=0.0, sparse_adam="lazy", device=dev, seed=999)
    net.load_state_dict({k: v.to(dev) for k, v in CK.load_state_dict(str(tmp_path / "cp")).items()})
    for name, t in zip(("uMLP", "iMLP", "uMF", "iMF"), sh._tables()):
        assert np.array_equal(getattr(net, name).w.cpu().numpy(), t.full_weights())
    assert torch.equal(net.dense.w.view(-1), sh.dense.w.view(-1)) and torch.equal(net.bn_moving, sh.bn_moving)
    sh2 = ShardedNeuMFNet(U, I, E, dropout=0.0, device=dev, mode="peer", emulate=2, seed=5)
    sh2.load_checkpoint(str(tmp_path / "cp"))
    for t, t2 in zip(sh._tables(), sh2._tables()):
        assert np.array_equal(t.full_weights(), t2.full_weights())
    assert torch.equal(sh2.dense.m, sh.dense.m) and torch.equal(sh2.optimizer.state, sh.optimizer.state)


def test_ncf_model_serves_like_the_reference(dev, tmp_path):
    """NCFModel (src/models/NCFModel.py): serve-only class around the script-spec network -- product / customer ids
    from the training file in first-appearance order (:60-64), checkpoint restore on a fresh object, predictForUser =
    predictions over ALL product ids, those >= 1.0 dropped, best numberOfItem as {product: '%.9f'} (:42-51)."""
    from binrec_b200.NCFModel import NCFModel
    csv = tmp_path / "test.csv"
    rng = np.random.default_rng(3)
    rows = ["customer_id,normalized_customer_id,material,product_id,rating_type"]
    cust = rng.integers(0, 40, 300); prod = rng.integers(0, 60, 300)
    rows += [f"{900000 + c},{c},{55000 + p},{p},{int(rng.random() < 0.5)}" for c, p in zip(cust, prod)]
    csv.write_text("\n".join(rows) + "\n")
    m = NCFModel(workDir=str(tmp_path), trainData=str(csv))
    assert m.productIds == list(dict.fromkeys(prod.tolist())) and m.customerIds == list(dict.fromkeys(cust.tolist()))
    assert not m.readyToTrain() and m.getPredictableUsers() == m.customerIds
    net = m.compileModel(None, 40, 60)
    assert (net.E, net.hidden, net.act, net.loss) == (10, (100, 50, 10), "sigmoid", "bce")
    u = torch.from_numpy(cust.astype(np.int32)).to(dev); i = torch.from_numpy(prod.astype(np.int32)).to(dev)
    y = torch.from_numpy((rng.random(300) < 0.5).astype(np.float32)).to(dev)
    for _ in range(5):
        net.train_on_batch(u, i, y)
    m.saveCheckPoint()
    fresh = NCFModel(workDir=str(tmp_path))
    fresh.restoreFromLatestCheckPoint()                                    # what the REST endpoint does
    assert fresh.productIds == m.productIds and fresh.getPredictableUsers() == m.customerIds
    customer = m.customerIds[3]
    frame = fresh.getPredictDataFrame(customer)
    assert frame["PRODUCT_ID"] == m.productIds and set(frame["CUSTOMER_ID"]) == {customer}
    pred, _ = net.predict_on_batch(torch.full((len(m.productIds),), customer, dtype=torch.int32, device=dev),
                                   torch.as_tensor(np.asarray(m.productIds, dtype=np.int32)).to(dev))
    pred = pred.cpu().numpy()
    want = dict(zip(m.productIds, pred))
    want = dict(filter(lambda e: e[1] < 1.0, want.items()))
    want = {k: '%.9f' % v for k, v in sorted(want.items(), key=lambda x: x[1], reverse=True)[:7]}
    got = fresh.predictForUser(customer, 7)
    assert list(got.items()) == list(want.items())
    many = fresh.predictForUsers(m.customerIds[:5], 3)
    assert [list(d) for d in many][3][:3] == list(want)[:3] and all(len(d) == 3 for d in many)
    with pytest.raises(ValueError):
        fresh.predictForUser(4000)
    with pytest.raises(ValueError):
        fresh.predictForUser(customer, 64)


def test_fit_nfc_plain_learns_a_planted_rule(dev):
    """fit_nfc_plain = the script trainers/NFC_plain.py as a function: explicit labels from the file, script-spec
    network, rows reshuffled per epoch; a planted parity rule must be learnt and the table sizes follow :79-82."""
    from binrec_b200.NCFModel import fit_nfc_plain
    rng = np.random.default_rng(0)
    C, P, n = 60, 90, 40000
    c = rng.integers(0, C, n); p = rng.integers(1, P + 1, n)
    y = ((c % 3 == 0) ^ (p % 2 == 0)).astype(np.float32)
    res = fit_nfc_plain((c[:32000], p[:32000], y[:32000]), (c[32000:], p[32000:], y[32000:]), epochs=30, batch_size=4000)
    net = res["model"]
    assert net.numUser == C + 1 and net.numItem == P + 1 and net.head_order == "mf_h3"
    assert res["history"][-1] < 0.5 * res["history"][0]
    loss, mse, mae, acc = res["evaluate"]
    assert acc > 0.9 and res["test_mae_rounded"] == pytest.approx(mae, abs=0.01)
