"""Generates tests/golden/wiring_golden.npz by EXECUTING the reference's own model-building and training-step code

    /root/reference/src/models/NeuMFModel.py:53-100     NeuMFModel.compileModel
    /root/reference/src/models/BPRModel.py:49-74,124-144  BPRModel.compileModel, bprTripletLoss, identityLoss
    /root/reference/trainers/twoTower.py:19-111         TwoTowerModel (constructor, computeEmb, computeLossTfrs,
                                                        computeLossRdZero, train_step, setCandidates, call)

against the torch-backed Keras / TFRS stand-in of tests/golden/keras_shim.py (TensorFlow cannot be installed here).
The reference builds the graphs, picks the losses and optimizers and drives the tape; the shim supplies the layer
arithmetic.  Weights are copied in from the oracles' seeded parameters (by layer creation order, after checking
the layer types and shapes the reference created), one training step is run in float64 and everything observable
is recorded: predictions / losses, every gradient, every weight after the reference's optimizer step, BatchNorm
moving statistics, top-k lists.  tests/test_oracle_wiring.py holds oracle/{neumf,bpr,twotower}.py to these files.

Runs only in the build container (it imports /root/reference); the .npz is committed.
    python tests/golden/make_wiring_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "wiring_golden.npz")

import keras_shim as K  # noqa: E402

K.install()
from oracle import neumf as ON  # noqa: E402
from oracle import bpr as OB  # noqa: E402
from oracle import twotower as OT  # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stub_rmodel():
    """src.models.RModel pulls in matplotlib, DataStore (reads c.json) and creates directories: the class surface the
    model files use is re-declared (attribute names and defaults of RModel.py:16-43,201-202)."""
    class RModel:
        CUSTOMER_ID = 'CUSTOMER_ID'
        PRODUCT_ID = 'PRODUCT_ID'
        METRICS = ['mse', 'mae', 'binary_accuracy']

        def __init__(self, moduleName):
            self.modelName = moduleName
            self.numFactor, self.epochs, self.batchSize, self.testSize = 32, 10, 1024, 0.2
            self.model = None

        def getNumberOfWorkers(self, distributedConfig):
            return len(distributedConfig['cluster']['worker'])

    rm = types.ModuleType("src.models.RModel"); rm.RModel = RModel
    for name, mod in (("src", types.ModuleType("src")), ("src.models", types.ModuleType("src.models")),
                      ("src.models.RModel", rm)):
        sys.modules[name] = mod


def set_(var, arr):
    with torch.no_grad():
        assert tuple(var.shape) == tuple(arr.shape), (tuple(var.shape), tuple(arr.shape))
        var.copy_(torch.as_tensor(np.asarray(arr), dtype=K.DT))


def neumf_case(out, tag, U, I, F, B, seed, dropout):
    """compileModel -> the shim's functional Model; weights from oracle/neumf.NeuMFParams; one training step."""
    K.reset()
    _stub_rmodel()
    ref = _load("src/models/NeuMFModel.py", "ref_neumf")
    m = ref.NeuMFModel()
    model = m.compileModel(None, U, I, F)
    kinds = [type(l).__name__ for l in model.layers]
    # what the reference built, in creation order (NeuMFModel.py:58-83)
    assert kinds == ["Embedding"] * 4 + ["Concatenate", "Dropout", "Dense", "BatchNormalization", "Dropout", "Dense",
                                         "BatchNormalization", "Dropout", "Dense", "Dot", "Concatenate", "Dense"], kinds
    assert isinstance(model.optimizer, K.Adam) and model.optimizer.lr == 1e-3 and model.loss == 'mean_squared_error'
    assert model.compiled_metrics_names == ['mse', 'mae', 'binary_accuracy']
    hidden = (F, F // 2, F // 4)
    p = ON.NeuMFParams(U, I, F, hidden, seed=seed, dtype=torch.float64)
    w = p.numpy()
    emb = [l for l in model.layers if isinstance(l, K.Embedding)]
    dense = [l for l in model.layers if isinstance(l, K.Dense)]
    bn = [l for l in model.layers if isinstance(l, K.BatchNormalization)]
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, B); i = rng.integers(0, I, B)
    y = (rng.random(B) < 0.3).astype(np.float64)
    feed = {"user": u.astype(np.float32), "item": i.astype(np.float32)}           # the reference feeds float ids
    model(feed, training=False)                                                     # builds Dense / BN variables
    for l, k in zip(emb, ("uMLP", "iMLP", "uMF", "iMF")):                            # creation order :58-63
        set_(l.embeddings, w[k])
    for l, (kw, kb) in zip(dense, (("W1", "b1"), ("W2", "b2"), ("W3", "b3"), ("W4", "b4"))):
        set_(l.kernel, w[kw]); set_(l.bias, w[kb])
    # non-trivial BatchNorm parameters so that gamma / beta wiring shows
    g1 = 1 + 0.1 * rng.standard_normal(hidden[0]); be1 = 0.1 * rng.standard_normal(hidden[0])
    g2 = 1 + 0.1 * rng.standard_normal(hidden[1]); be2 = 0.1 * rng.standard_normal(hidden[1])
    for l, (g, b) in zip(bn, ((g1, be1), (g2, be2))):
        set_(l.gamma, g); set_(l.beta, b)
    first, epoch, dseed = 4096, 3, 11
    idx = np.arange(first, first + B, dtype=np.uint64)
    masks = None
    if dropout:
        masks = [ON.dropout_mask(n, idx, L, dseed, epoch) for L, n in ((0, 2 * F), (1, hidden[0]), (2, hidden[1]))]
        for k in range(3):
            K.DROPOUT_MASKS[k] = torch.as_tensor(masks[k], dtype=K.DT)
    pred = model(feed, training=True)
    loss = model.compiled_loss(y, pred)
    names = ("uMLP", "iMLP", "uMF", "iMF", "W1", "b1", "g1", "be1", "W2", "b2", "g2", "be2", "W3", "b3", "W4", "b4")
    variables = [emb[0].embeddings, emb[1].embeddings, emb[2].embeddings, emb[3].embeddings, dense[0].kernel, dense[0].bias,
                 bn[0].gamma, bn[0].beta, dense[1].kernel, dense[1].bias, bn[1].gamma, bn[1].beta, dense[2].kernel,
                 dense[2].bias, dense[3].kernel, dense[3].bias]
    assert {id(v) for v in variables} == {id(v) for v in model.trainable_variables}
    grads = K.GradientTape().gradient(loss, variables)
    model.optimizer.apply_gradients(zip(grads, variables))
    infer = model(feed, training=False)
    out.update({f"{tag}/u": u, f"{tag}/i": i, f"{tag}/y": y, f"{tag}/meta": np.array([U, I, F, seed, first, epoch, dseed, int(dropout)]),
                f"{tag}/g1": g1, f"{tag}/be1": be1, f"{tag}/g2": g2, f"{tag}/be2": be2,
                f"{tag}/pred": pred.detach().numpy().reshape(-1), f"{tag}/loss": np.array(float(loss.detach())),
                f"{tag}/infer": infer.detach().numpy().reshape(-1),
                f"{tag}/mm1": bn[0].moving_mean.numpy(), f"{tag}/mv1": bn[0].moving_variance.numpy(),
                f"{tag}/mm2": bn[1].moving_mean.numpy(), f"{tag}/mv2": bn[1].moving_variance.numpy()})
    for n, g, v in zip(names, grads, variables):
        out[f"{tag}/grad/{n}"] = g.detach().numpy()
        out[f"{tag}/after/{n}"] = v.detach().numpy()
    print(tag, "loss", float(loss.detach()))


def bpr_case(out, tag, U, I, d, B, seed):
    K.reset()
    _stub_rmodel()
    ref = _load("src/models/BPRModel.py", "ref_bpr")
    m = ref.BPRModel()
    model, strategy = m.compileModel(None, U, I, d)
    assert strategy is None
    kinds = [type(l).__name__ for l in model.layers]
    # item table first (shared by positive and negative), then the user table (BPRModel.py:55-61)
    assert kinds == ["Embedding", "Flatten", "Flatten", "Embedding", "Flatten", "Lambda"], kinds
    assert [l.name for l in model.layers if isinstance(l, K.Embedding)] == ["item_embedding", "user_embedding"]
    assert isinstance(model.optimizer, K.Adam) and model.optimizer.lr == 1e-3
    orc = OB.BPROracle(U, I, d, seed=seed)
    item_l, user_l = [l for l in model.layers if isinstance(l, K.Embedding)]
    set_(user_l.embeddings, orc.user); set_(item_l.embeddings, orc.item)
    rng = np.random.default_rng(seed + 1)
    u = rng.integers(0, U, B); p = rng.integers(0, I, B); n = rng.integers(0, I, B)
    feed = {"customerId_input": u.astype(np.float32).reshape(B, 1), "pProduct_input": p.astype(np.float32).reshape(B, 1),
            "nProduct_input": n.astype(np.float32).reshape(B, 1)}                   # BPRModel.py:100-104
    y_pred = model(feed, training=True)
    assert tuple(y_pred.shape) == (B, 1)
    loss = model.compiled_loss(torch.ones(B), y_pred)                               # identityLoss(_, y_pred), :121-122
    variables = [user_l.embeddings, item_l.embeddings]
    grads = K.GradientTape().gradient(loss, variables)
    model.optimizer.apply_gradients(zip(grads, variables))
    out.update({f"{tag}/u": u, f"{tag}/p": p, f"{tag}/n": n, f"{tag}/meta": np.array([U, I, d, seed]),
                f"{tag}/triplet": y_pred.detach().numpy().reshape(-1), f"{tag}/loss": np.array(float(loss.detach())),
                f"{tag}/grad/user": grads[0].numpy(), f"{tag}/grad/item": grads[1].numpy(),
                f"{tag}/after/user": user_l.embeddings.detach().numpy(), f"{tag}/after/item": item_l.embeddings.detach().numpy()})
    print(tag, "loss", float(loss.detach()))


def twotower_case(out, tag, U, I, E, S, B, seed, rdZero):
    K.reset()
    # modules twoTower.py imports at the top that are absent or irrelevant here
    for name, attrs in (("src.benchmarkLogger", dict(benchThread=object)), ("trainers.model_utils", dict(getOptimizer=None)),
                        ("trainers.loadBinaryMovieLens", {}), ("trainers.topKmetrics", {})):
        mod = types.ModuleType(name); mod.__dict__.update(attrs); sys.modules[name] = mod
    for name in ("src", "trainers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ref = _load("trainers/twoTower.py", "ref_twotower")
    users = [f"u{j}" for j in range(U)]; items = [f"m{j}" for j in range(I)]
    model = ref.TwoTowerModel(E, I, U, "CUSTOMER_ID", "MATERIAL", users, items, eval_batch_size=64, loss=None,
                              rdZero=rdZero, resKey="RATING_TYPE", semb=S)
    # crossValidation's compile call (twoTower.py:209) with the Adagrad(0.1) of the CLI defaults (:278-279)
    model.compile(optimizer=K.Adagrad(learning_rate=0.1), loss=K.BinaryCrossentropy())
    orc = OT.TwoTowerOracle(U, I, E, S, seed=seed, dtype=torch.float64, rdZero=rdZero)
    w = {k: v.detach().numpy() for k, v in orc.t.items()}
    rng = np.random.default_rng(seed + 2)
    ui = rng.integers(0, U, B); ii = rng.integers(0, max(I // 3, 2), B)             # repeated items: accidental hits occur
    labels = (rng.random(B) < 0.5).astype(np.float64)
    info = {"CUSTOMER_ID": np.array([users[j] for j in ui]), "MATERIAL": np.array([items[j] for j in ii]),
            "RATING_TYPE": labels}
    model.computeEmb(info)                                                          # builds the Dense kernels
    ut, it = model.userTower.layers, model.itemTower.layers
    assert [type(l).__name__ for l in ut] == ["StringLookup", "Embedding", "Dense"]
    assert tuple(ut[1].embeddings.shape) == (U + 2, E) and tuple(it[1].embeddings.shape) == (I + 2, E)
    assert ut[2].activation is None and ut[2].units == S
    set_(ut[1].embeddings, w["Eu"]); set_(it[1].embeddings, w["Ei"])
    set_(ut[2].kernel, w["Wu"]); set_(it[2].kernel, w["Wi"])
    bu = 0.05 * rng.standard_normal(S); bi = 0.05 * rng.standard_normal(S)
    set_(ut[2].bias, bu); set_(it[2].bias, bi)
    variables = model.trainable_variables
    # attribute order of the constructor (twoTower.py:33-41): both Embedding layers, then the towers' Dense layers
    order = ["Eu", "Ei", "Wu", "bu", "Wi", "bi"]
    assert [tuple(v.shape) for v in variables] == [(U + 2, E), (I + 2, E), (E, S), (S,), (E, S), (S,)]
    before = [v.detach().clone() for v in variables]
    metrics = model.train_step(info)                                                # tape, gradient, apply_gradients :89-102
    loss = float(metrics["loss"].detach())
    if not rdZero:
        assert model.task.calls == [dict(compute_metrics=False, training=True, has_ids=True)]
    # gradients of the same loss at the weights before the step (train_step does not return them)
    with torch.no_grad():
        after = [v.detach().clone() for v in variables]
        for v, b in zip(variables, before):
            v.copy_(b)
    q, c = model.computeEmb(info)
    l2 = model.computeLoss(q, c, info)
    grads = K.GradientTape().gradient(l2, variables)
    assert abs(float(l2.detach()) - loss) < 1e-12
    with torch.no_grad():
        for v, a in zip(variables, after):
            v.copy_(a)
    out.update({f"{tag}/ui": ui, f"{tag}/ii": ii, f"{tag}/labels": labels, f"{tag}/bu": bu, f"{tag}/bi": bi,
                f"{tag}/meta": np.array([U, I, E, S, seed, int(rdZero)]), f"{tag}/loss": np.array(loss)})
    for n, g, a in zip(order, grads, after):
        out[f"{tag}/grad/{n}"] = g.numpy(); out[f"{tag}/after/{n}"] = a.numpy()
    if not rdZero:
        # evaluation path: setCandidates(items, k) + call(users) (twoTower.py:60-69, 229-230)
        k = 5
        model.setCandidates(K.Dataset.from_tensor_slices(items), k)
        vals, ids = model.call(np.array(users))
        out[f"{tag}/topk_vals"] = vals.detach().numpy(); out[f"{tag}/topk_ids"] = np.array([[int(s[1:]) for s in r] for r in ids])
    print(tag, "loss", loss)


def main():
    torch.manual_seed(0)
    out = {}
    neumf_case(out, "neumf_f32", 60, 40, 32, 96, 42, dropout=True)
    neumf_case(out, "neumf_f8_nodrop", 30, 20, 8, 50, 7, dropout=False)
    neumf_case(out, "neumf_f20", 25, 35, 20, 64, 3, dropout=True)               # numFactor is a free attribute (RModel.py:35)
    bpr_case(out, "bpr_d64", 50, 30, 64, 80, 42)
    bpr_case(out, "bpr_d350", 20, 25, 350, 40, 5)                               # bpr.py:21 latent 350
    twotower_case(out, "tt_tfrs", 40, 30, 16, 12, 48, 42, rdZero=False)
    twotower_case(out, "tt_rdzero", 40, 30, 16, 12, 48, 9, rdZero=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
