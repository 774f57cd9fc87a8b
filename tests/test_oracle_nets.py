"""Internal consistency of the two autograd oracles that cannot be pinned against the reference (TensorFlow / Keras /
TFRS are not installable here): an independent NumPy restatement of the forward pass written
from the reference's layer list, and central finite differences of the loss in float64 for the gradients.
  NeuMF     : src/models/NeuMFModel.py:53-100  (BatchNorm AFTER the activation, scalar Dot, head [h3, mf])
  two-tower : trainers/twoTower.py:19-111      (linear towers, in-batch softmax SUM-reduced with accidental-hit removal)"""
import numpy as np
import torch

from oracle import neumf as ON
from oracle import twotower as OT


def _np_neumf(P, u, i, act):
    """Training-mode forward in plain NumPy float64 (no dropout): the layer list of NeuMFModel.py:58-83."""
    f = (lambda x: np.maximum(x, 0.0)) if act == "relu" else (lambda x: 1.0 / (1.0 + np.exp(-x)))
    def bn(h, g, b):
        mu, var = h.mean(0), h.var(0)                                     # biased batch variance (Keras)
        return g * (h - mu) / np.sqrt(var + 1e-3) + b
    x0 = np.concatenate([P["uMLP"][u], P["iMLP"][i]], axis=1)             # concat[uMLP, iMLP]            :66
    y1 = bn(f(x0 @ P["W1"] + P["b1"]), P["g1"], P["be1"])                 # Dense(F, act) -> BN           :69-70
    y2 = bn(f(y1 @ P["W2"] + P["b2"]), P["g2"], P["be2"])                 # Dense(F/2, act) -> BN         :73-74
    h3 = f(y2 @ P["W3"] + P["b3"])                                        # Dense(F/4, act)               :78
    mf = (P["uMF"][u] * P["iMF"][i]).sum(1, keepdims=True)                # Dot(axes=1): a scalar         :79
    logit = (np.concatenate([h3, mf], axis=1) @ P["W4"] + P["b4"])[:, 0]  # concat[h3, predMF] -> Dense(1):80-83
    return 1.0 / (1.0 + np.exp(-logit)), logit


def test_neumf_oracle_forward_equals_the_numpy_restatement_and_grads_match_finite_differences():
    rng = np.random.default_rng(0)
    for act, loss in (("relu", "mse"), ("sigmoid", "bce")):
        p = ON.NeuMFParams(7, 5, 4, (6, 4, 2), seed=3, dtype=torch.float64)
        with torch.no_grad():                                             # leave the all-zero biases of the init
            for k in ("b1", "b2", "b3", "b4", "be1", "be2"):
                p.t[k] += torch.from_numpy(rng.normal(0, 0.1, p.t[k].shape))
        u = rng.integers(0, 7, 9); i = rng.integers(0, 5, 9); y = (rng.random(9) < 0.4).astype(np.float64)
        out, aux = ON.forward(p, u, i, training=True, act=act)
        ref_out, ref_logit = _np_neumf(p.numpy(), u, i, act)
        np.testing.assert_allclose(out.detach().numpy(), ref_out, rtol=1e-12)
        np.testing.assert_allclose(aux["logit"].detach().numpy(), ref_logit, rtol=1e-10, atol=1e-13)

        def loss_np(P):
            o, lg = _np_neumf(P, u, i, act)
            if loss == "mse":
                return float(((o - y) ** 2).mean())
            return float(np.mean(np.maximum(lg, 0) - lg * y + np.log1p(np.exp(-np.abs(lg)))))   # BCE from logits

        l = ON.loss_fn(out, aux["logit"], torch.from_numpy(y), loss)
        np.testing.assert_allclose(float(l.detach()), loss_np(p.numpy()), rtol=1e-12)
        l.backward()
        base = p.numpy()
        for name in ("uMLP", "iMF", "W1", "g1", "be2", "W3", "W4", "b4"):
            g = p.t[name].grad.numpy()
            flat = np.argsort(-np.abs(g).ravel())[:3]                     # the three largest entries
            for idx in flat:
                pos = np.unravel_index(idx, g.shape)
                hi = {k: v.copy() for k, v in base.items()}; lo = {k: v.copy() for k, v in base.items()}
                hi[name][pos] += 1e-6; lo[name][pos] -= 1e-6
                fd = (loss_np(hi) - loss_np(lo)) / 2e-6
                np.testing.assert_allclose(g[pos], fd, rtol=2e-5, atol=1e-9, err_msg=f"{act} {name}{pos}")


def _np_twotower_loss(T, u, i, cand):
    q = T["Eu"][u] @ T["Wu"] + T["bu"]; c = T["Ei"][i] @ T["Wi"] + T["bi"]       # linear towers  :40-41
    s = q @ c.T
    dup = (cand[:, None] == cand[None, :]) & ~np.eye(len(cand), dtype=bool)      # accidental hits
    s = s + dup * (float(np.finfo(np.float32).min) / 100.0)
    s = s - s.max(1, keepdims=True)
    return float(-(np.diag(s) - np.log(np.exp(s).sum(1))).sum())                 # categorical CE, reduction SUM


def test_twotower_oracle_loss_equals_the_numpy_restatement_and_grads_match_finite_differences():
    rng = np.random.default_rng(1)
    o = OT.TwoTowerOracle(9, 6, 5, 3, seed=2, dtype=torch.float64)
    u = rng.integers(2, 11, 12); i = rng.integers(2, 8, 12)                       # StringLookup offset 2; repeated items
    assert len(set(i.tolist())) < 12
    l = o.loss_and_grads(u, i, cand_ids=i)
    base = {k: v.detach().numpy().copy() for k, v in o.t.items()}
    np.testing.assert_allclose(float(l), _np_twotower_loss(base, u, i, i), rtol=1e-12)
    for name in ("Eu", "Ei", "Wu", "Wi", "bu", "bi"):
        g = o.t[name].grad.numpy()
        for idx in np.argsort(-np.abs(g).ravel())[:3]:
            pos = np.unravel_index(idx, g.shape)
            hi = {k: v.copy() for k, v in base.items()}; lo = {k: v.copy() for k, v in base.items()}
            hi[name][pos] += 1e-6; lo[name][pos] -= 1e-6
            fd = (_np_twotower_loss(hi, u, i, i) - _np_twotower_loss(lo, u, i, i)) / 2e-6
            np.testing.assert_allclose(g[pos], fd, rtol=2e-5, atol=1e-8, err_msg=f"{name}{pos}")
    # one Adagrad step: accumulator 0.1 + g^2, w -= lr g / (sqrt(acc) + 1e-7)                      :278-279
    g = {k: v.grad.numpy().copy() for k, v in o.t.items()}
    o.step(u, i, cand_ids=i)
    for k in base:
        want = base[k] - 0.1 * g[k] / (np.sqrt(0.1 + g[k] ** 2) + 1e-7)
        np.testing.assert_allclose(o.t[k].detach().numpy(), want, rtol=1e-12, atol=1e-15)
