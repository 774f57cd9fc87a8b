"""The oracles against tests/golden/wiring_golden.npz: outputs of the reference's OWN graph-building and training-step
code (NeuMFModel.compileModel, BPRModel.compileModel + bprTripletLoss + identityLoss, TwoTowerModel incl. train_step,
setCandidates and call) executed over the torch-backed Keras / TFRS stand-in of tests/golden/keras_shim.py by
tests/golden/make_wiring_golden.py.  This pins the WIRING of oracle/neumf.py, oracle/bpr.py and oracle/twotower.py to
executed reference code; the arithmetic inside each Keras / TFRS layer stays "upstream numerics, restated".

float64 on both sides, so the tolerance is summation-order noise: rtol 1e-9 / atol 1e-12."""
import os

import numpy as np
import pytest
import torch

from oracle import bpr as OB
from oracle import embedding as OE
from oracle import neumf as ON
from oracle import twotower as OT

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wiring_golden.npz"))
TOL = dict(rtol=1e-9, atol=1e-12)


def _g(tag, name):
    return G[f"{tag}/{name}"]


@pytest.mark.parametrize("tag", ["neumf_f32", "neumf_f8_nodrop", "neumf_f20"])
def test_neumf_oracle_matches_executed_compileModel(tag):
    U, I, F, seed, first, epoch, dseed, dropout = (int(x) for x in _g(tag, "meta"))
    orc = ON.NeuMFOracle(U, I, emb=F, seed=seed, dropout=0.2 if dropout else 0.0, dropout_seed=dseed, dtype=torch.float64)
    with torch.no_grad():
        for k in ("g1", "be1", "g2", "be2"):
            orc.p.t[k].copy_(torch.as_tensor(_g(tag, k)))
    u, i, y = _g(tag, "u"), _g(tag, "i"), _g(tag, "y")
    loss, out, aux = orc.loss_and_grads(u, i, y, first_index=first, epoch=epoch)
    np.testing.assert_allclose(out.numpy(), _g(tag, "pred"), **TOL)
    np.testing.assert_allclose(float(loss), float(_g(tag, "loss")), **TOL)
    for name in ON.NeuMFParams.TABLES + ON.NeuMFParams.DENSE:
        np.testing.assert_allclose(orc.p.t[name].grad.numpy(), _g(tag, f"grad/{name}"), err_msg=name, **TOL)
    orc.step(u, i, y, first_index=first, epoch=epoch)                       # Adam(1e-3) as compiled by the reference
    for name in ON.NeuMFParams.TABLES + ON.NeuMFParams.DENSE:
        np.testing.assert_allclose(orc.p.t[name].detach().numpy(), _g(tag, f"after/{name}"), err_msg=name, **TOL)
    for mine, name in ((orc.p.mm1, "mm1"), (orc.p.mv1, "mv1"), (orc.p.mm2, "mm2"), (orc.p.mv2, "mv2")):
        np.testing.assert_allclose(mine.numpy(), _g(tag, name), err_msg=name, **TOL)
    np.testing.assert_allclose(orc.predict(u, i), _g(tag, "infer"), **TOL)  # inference: moving statistics, no dropout


@pytest.mark.parametrize("tag", ["bpr_d64", "bpr_d350"])
def test_bpr_oracle_matches_executed_triplet_graph(tag):
    U, I, d, seed = (int(x) for x in _g(tag, "meta"))
    orc = OB.BPROracle(U, I, d, seed=seed)                                  # float32 initial tables (as the golden run), float64 arithmetic
    for k in ("user", "item", "mu", "vu", "mi", "vi"):
        setattr(orc, k, getattr(orc, k).astype(np.float64))
    u, p, n = _g(tag, "u"), _g(tag, "p"), _g(tag, "n")
    x, s = OB.bpr_forward(orc.user, orc.item, u, p, n)
    np.testing.assert_allclose(1.0 - s, _g(tag, "triplet"), **TOL)          # bprTripletLoss output per triplet
    loss, gu, gi = OB.bpr_loss_and_grads(orc.user, orc.item, u, p, n)
    np.testing.assert_allclose(float(loss), float(_g(tag, "loss")), **TOL)  # identityLoss = mean
    np.testing.assert_allclose(gu, _g(tag, "grad/user"), **TOL)
    np.testing.assert_allclose(gi, _g(tag, "grad/item"), **TOL)             # ONE item table shared by positive and negative
    orc.step(u, p, n)
    np.testing.assert_allclose(orc.user, _g(tag, "after/user"), **TOL)
    np.testing.assert_allclose(orc.item, _g(tag, "after/item"), **TOL)


@pytest.mark.parametrize("tag", ["tt_tfrs", "tt_rdzero"])
def test_twotower_oracle_matches_executed_train_step(tag):
    U, I, E, S, seed, rdZero = (int(x) for x in _g(tag, "meta"))
    orc = OT.TwoTowerOracle(U, I, E, S, seed=seed, dtype=torch.float64, rdZero=bool(rdZero))
    with torch.no_grad():
        orc.t["bu"].copy_(torch.as_tensor(_g(tag, "bu"))); orc.t["bi"].copy_(torch.as_tensor(_g(tag, "bi")))
    ui, ii, labels = _g(tag, "ui"), _g(tag, "ii"), _g(tag, "labels")
    # StringLookup: vocabulary entry j -> index j + 2; the reference passes the MATERIAL column as candidate_ids
    loss = orc.loss_and_grads(ui + 2, ii + 2, cand_ids=None if rdZero else ii, labels=labels if rdZero else None)
    np.testing.assert_allclose(float(loss), float(_g(tag, "loss")), **TOL)
    for name in ("Eu", "Ei", "Wu", "bu", "Wi", "bi"):
        np.testing.assert_allclose(orc.t[name].grad.numpy(), _g(tag, f"grad/{name}"), err_msg=name, **TOL)
    orc.step(ui + 2, ii + 2, cand_ids=None if rdZero else ii, labels=labels if rdZero else None)   # Adagrad(0.1)
    for name in ("Eu", "Ei", "Wu", "bu", "Wi", "bi"):
        np.testing.assert_allclose(orc.t[name].detach().numpy(), _g(tag, f"after/{name}"), err_msg=name, **TOL)
    if not rdZero:
        # BruteForce evaluation path (setCandidates + call): scores = Q C^T, top-k sorted, ties -> lower index
        from oracle import topk as OK
        q = orc.user_vectors(np.arange(U) + 2); c = orc.item_vectors(np.arange(I) + 2)
        vals, ids = OK.topk_from_scores(q @ c.T, _g(tag, "topk_ids").shape[1])
        assert np.array_equal(ids, _g(tag, "topk_ids"))
        np.testing.assert_allclose(vals, _g(tag, "topk_vals"), **TOL)
