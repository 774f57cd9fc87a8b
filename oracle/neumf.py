"""ORACLE (test infrastructure, never the product path): NeuMF restated on the CPU with torch
autograd (fp32 or fp64).

Pinning.  WIRING PINNED: the reference's own NeuMFModel.compileModel is EXECUTED over a torch-backed Keras stand-in
(tests/golden/keras_shim.py, tests/golden/make_wiring_golden.py -> tests/golden/wiring_golden.npz) and this oracle
reproduces its predictions, loss, every gradient, the weights after the compiled Adam(1e-3) step, the BatchNorm moving
statistics and the inference output to 1e-9 in float64 (tests/test_oracle_wiring.py; numFactor 32, 8 and 20).
UPSTREAM NUMERICS UNPINNED: the arithmetic INSIDE each Keras layer (Dense, BatchNormalization, Dropout, Adam ...) is
restated from the upstream documentation in both the shim and here -- TensorFlow cannot be installed and the reference
holds no golden vectors -- and is held to an independent NumPy restatement and fp64 finite
differences (tests/test_oracle_nets.py).  The He et al. variant below is not in the reference tree at all.

Follows /root/reference/src/models/NeuMFModel.py:53-100 (class spec) and
/root/reference/trainers/NFC_plain.py:109-155 (script spec):

  x0   = concat[uMLP[u], iMLP[i]]               NeuMFModel.py:58-66   (script: [item, user], NFC_plain.py:137)
  d0   = dropout(x0, 0.2)                        :67
  h1   = act(d0 W1 + b1); y1 = BN(h1); d1 = dropout(y1)     :69-71    (BN AFTER the activation)
  h2   = act(d1 W2 + b2); y2 = BN(h2); d2 = dropout(y2)     :73-75
  h3   = act(d2 W3 + b3)                          :78
  mf   = sum_f uMF[u,f] * iMF[i,f]                :79   (Dot(axes=1): a scalar, not a Hadamard vector)
  out  = sigmoid(concat[h3, mf] W4 + b4)          :80-83 (script: concat[mf, h3], NFC_plain.py:149)
  loss = mean((out - y)^2)  (class)  |  BCE (script, NFC_plain.py:155)

The head weight is stored here in the class order [h3..., mf]; the script order is a row permutation
of W4 handled by the host wrapper.  Keras defaults: BN momentum 0.99, eps 1e-3, biased batch variance,
gamma 1, beta 0, moving mean 0 / variance 1; inverted dropout; Dense glorot-uniform / zero bias;
Embedding U(-0.05, 0.05).

Dropout masks are not reproducible from TensorFlow; they are DEFINED here from Philox so that the
device can regenerate them (stream tags 0xD0 + layer): one Philox call yields 16 bytes, feature
4*c16+.. keeps iff its byte >= 51, i.e. keep probability 205/256 (0.8008 instead of 0.8) with scale
256/205.  rate == 0 disables dropout entirely.

Variant (BASELINE.json configs[0], "NeuMF (GMF+MLP, 8-dim, layers 64-32-16-8)", He et al. 2017 -- NOT in the
reference tree, defined here): mf_mode="hadamard" concatenates the element-wise product uMF[u]*iMF[i] [mf_dim]
to h3 (W4 has H3 + mf_dim rows) instead of the scalar Dot; batch_norm=False drops both BatchNormalization layers.

matmul: the Dense products go through `matmul(x, W)`; oracle/tf32.py supplies the TF32-operand form the tensor-core
kernels compute (csrc/neumf_tc.cu, csrc/neumf_fused.cu).
"""
import numpy as np
import torch

from . import philox as P

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
DROP_THRESHOLD = 51            # byte < 51 -> dropped
DROP_KEEP = (256 - DROP_THRESHOLD) / 256.0


def dropout_mask(n_features, sample_index, layer, seed, epoch):
    """float mask [N, n_features] in {0, 1/keep}.  sample_index: uint64 array of global sample ids."""
    idx = np.asarray(sample_index, dtype=np.uint64)
    n = len(idx)
    calls = (n_features + 15) // 16
    out = np.empty((n, calls * 16), dtype=np.float32)
    for c in range(calls):
        ctr = np.stack([idx & P.MASK, idx >> np.uint64(32), np.full(n, c, dtype=np.uint64),
                        np.full(n, 0xD0 + layer, dtype=np.uint64)], axis=1).astype(np.uint32)
        w = P.philox4x32_10(ctr, (seed, epoch))                      # [n, 4] uint32
        b = w.view(np.uint8).reshape(n, 16)                          # little-endian bytes of words 0..3
        out[:, c * 16:(c + 1) * 16] = (b >= DROP_THRESHOLD).astype(np.float32) / np.float32(DROP_KEEP)
    return out[:, :n_features]


class NeuMFParams:
    """Plain container of torch tensors (requires_grad) in Keras layouts (Dense kernel [in, out])."""

    def __init__(self, num_users, num_items, emb, hidden, seed=42, dtype=torch.float32, mf_dim=None, mf_mode="dot",
                 batch_norm=True):
        rng = np.random.Generator(np.random.Philox(key=seed))
        h1, h2, h3 = hidden
        npdt = np.float64 if dtype == torch.float64 else np.float32
        mf_dim = emb if mf_dim is None else int(mf_dim)
        head_mf = mf_dim if mf_mode == "hadamard" else 1
        self.mf_dim, self.mf_mode, self.batch_norm = mf_dim, mf_mode, bool(batch_norm)

        def emb_init(rows, width=emb):
            return rng.uniform(-0.05, 0.05, size=(rows, width)).astype(np.float32)

        def glorot(i, o):
            lim = np.sqrt(6.0 / (i + o))
            return rng.uniform(-lim, lim, size=(i, o)).astype(np.float32)

        arrs = dict(uMLP=emb_init(num_users), iMLP=emb_init(num_items), uMF=emb_init(num_users, mf_dim),
                    iMF=emb_init(num_items, mf_dim), W1=glorot(2 * emb, h1), b1=np.zeros(h1, np.float32),
                    W2=glorot(h1, h2), b2=np.zeros(h2, np.float32), W3=glorot(h2, h3), b3=np.zeros(h3, np.float32),
                    W4=glorot(h3 + head_mf, 1), b4=np.zeros(1, np.float32),
                    g1=np.ones(h1, np.float32), be1=np.zeros(h1, np.float32),
                    g2=np.ones(h2, np.float32), be2=np.zeros(h2, np.float32))
        self.t = {k: torch.tensor(v.astype(npdt), requires_grad=True) for k, v in arrs.items()}
        self.mm1 = torch.zeros(h1, dtype=dtype); self.mv1 = torch.ones(h1, dtype=dtype)
        self.mm2 = torch.zeros(h2, dtype=dtype); self.mv2 = torch.ones(h2, dtype=dtype)
        self.emb, self.hidden = emb, hidden

    TABLES = ("uMLP", "iMLP", "uMF", "iMF")
    DENSE = ("W1", "b1", "g1", "be1", "W2", "b2", "g2", "be2", "W3", "b3", "W4", "b4")

    def numpy(self):
        return {k: v.detach().numpy().copy() for k, v in self.t.items()}


def _act(x, kind):
    return torch.relu(x) if kind == "relu" else torch.sigmoid(x)


def _bn(h, gamma, beta, training, mm, mv):
    if training:
        mu = h.mean(0); var = h.var(0, unbiased=False)
    else:
        mu, var = mm, mv
    return gamma * (h - mu) / torch.sqrt(var + BN_EPS) + beta, mu, var


def forward(p, u, i, training=False, act="relu", masks=None, matmul=torch.matmul):
    """Returns (out [B], aux dict).  masks: None or (m0 [B,2E], m1 [B,H1], m2 [B,H2]) float masks."""
    t = p.t
    u = torch.as_tensor(np.asarray(u), dtype=torch.int64); i = torch.as_tensor(np.asarray(i), dtype=torch.int64)
    bn_on = getattr(p, "batch_norm", True)

    def bn(h, g, be, mm, mv):
        if not bn_on:
            return h, mm, mv
        return _bn(h, g, be, training, mm, mv)

    x0 = torch.cat([t["uMLP"][u], t["iMLP"][i]], dim=1)
    if masks is not None:
        x0 = x0 * masks[0]
    h1 = _act(matmul(x0, t["W1"]) + t["b1"], act)
    y1, mu1, var1 = bn(h1, t["g1"], t["be1"], p.mm1, p.mv1)
    if masks is not None:
        y1 = y1 * masks[1]
    h2 = _act(matmul(y1, t["W2"]) + t["b2"], act)
    y2, mu2, var2 = bn(h2, t["g2"], t["be2"], p.mm2, p.mv2)
    if masks is not None:
        y2 = y2 * masks[2]
    h3 = _act(matmul(y2, t["W3"]) + t["b3"], act)
    if getattr(p, "mf_mode", "dot") == "hadamard":
        mf = t["uMF"][u] * t["iMF"][i]                       # GMF vector of He et al.
    else:
        mf = (t["uMF"][u] * t["iMF"][i]).sum(1, keepdim=True)
    logit = (torch.cat([h3, mf], dim=1) @ t["W4"] + t["b4"]).squeeze(1)
    return torch.sigmoid(logit), dict(logit=logit, mu1=mu1, var1=var1, mu2=mu2, var2=var2)


def loss_fn(out, logit, y, kind):
    if kind == "mse":
        return ((out - y) ** 2).mean()
    # Keras BinaryCrossentropy on a sigmoid output back-computes the logits; from-logits form
    return torch.nn.functional.binary_cross_entropy_with_logits(logit, y)


def metrics(out, y):
    """Keras metrics list ['mse', 'mae', 'binary_accuracy'] (RModel.py:20), threshold 0.5."""
    return {"mse": float(((out - y) ** 2).mean()), "mae": float((out - y).abs().mean()),
            "binary_accuracy": float(((out > 0.5).to(y.dtype) == y).to(y.dtype).mean())}


class NeuMFOracle:
    """Training-loop state with Keras Adam (dense-equivalent on the embedding tables)."""

    def __init__(self, num_users, num_items, emb=32, hidden=None, seed=42, act="relu", loss="mse", lr=1e-3,
                 dropout=0.0, dropout_seed=11, dtype=torch.float32, lazy_adam=False, mf_dim=None, mf_mode="dot",
                 batch_norm=True, matmul=torch.matmul):
        hidden = hidden or (emb, emb // 2, emb // 4)
        self.p = NeuMFParams(num_users, num_items, emb, hidden, seed, dtype, mf_dim, mf_mode, batch_norm)
        self.matmul = matmul
        self.act, self.loss, self.lr, self.dropout, self.dropout_seed = act, loss, lr, dropout, dropout_seed
        self.dtype = dtype
        self.m = {k: torch.zeros_like(v) for k, v in self.p.t.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.t.items()}
        self.t = 0
        self.lazy = lazy_adam

    def _masks(self, first_index, B, epoch):
        if self.dropout <= 0:
            return None
        idx = np.arange(first_index, first_index + B, dtype=np.uint64)
        E, (h1, h2, _) = self.p.emb, self.p.hidden
        return tuple(torch.from_numpy(dropout_mask(n, idx, L, self.dropout_seed, epoch)).to(self.dtype)
                     for L, n in ((0, 2 * E), (1, h1), (2, h2)))

    def loss_and_grads(self, u, i, y, first_index=0, epoch=0):
        for v in self.p.t.values():
            v.grad = None
        y = torch.as_tensor(np.asarray(y), dtype=self.dtype)
        out, aux = forward(self.p, u, i, training=True, act=self.act, masks=self._masks(first_index, len(y), epoch),
                           matmul=self.matmul)
        loss = loss_fn(out, aux["logit"], y, self.loss)
        loss.backward()
        return loss.detach(), out.detach(), aux

    def step(self, u, i, y, first_index=0, epoch=0):
        loss, out, aux = self.loss_and_grads(u, i, y, first_index, epoch)
        B = len(np.asarray(y))
        with torch.no_grad():
            # Keras BN moving statistics: moving = moving*momentum + batch*(1-momentum)  (biased variance)
            if self.p.batch_norm:
                self.p.mm1.mul_(BN_MOMENTUM).add_(aux["mu1"].detach() * (1 - BN_MOMENTUM))
                self.p.mv1.mul_(BN_MOMENTUM).add_(aux["var1"].detach() * (1 - BN_MOMENTUM))
                self.p.mm2.mul_(BN_MOMENTUM).add_(aux["mu2"].detach() * (1 - BN_MOMENTUM))
                self.p.mv2.mul_(BN_MOMENTUM).add_(aux["var2"].detach() * (1 - BN_MOMENTUM))
            self.t += 1
            b1, b2, eps = 0.9, 0.999, 1e-7
            alpha = self.lr * np.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
            uu = torch.as_tensor(np.unique(np.asarray(u)), dtype=torch.int64)
            ii = torch.as_tensor(np.unique(np.asarray(i)), dtype=torch.int64)
            for k, w in self.p.t.items():
                g = w.grad if w.grad is not None else torch.zeros_like(w)
                if self.lazy and k in NeuMFParams.TABLES:
                    rows = uu if k[0] == "u" else ii
                    self.m[k][rows] = b1 * self.m[k][rows] + (1 - b1) * g[rows]
                    self.v[k][rows] = b2 * self.v[k][rows] + (1 - b2) * g[rows] ** 2
                    w[rows] -= alpha * self.m[k][rows] / (self.v[k][rows].sqrt() + eps)
                else:
                    self.m[k].mul_(b1).add_(g, alpha=1 - b1)
                    self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
                    w.sub_(alpha * self.m[k] / (self.v[k].sqrt() + eps))
        return float(loss), metrics(out, torch.as_tensor(np.asarray(y), dtype=self.dtype))

    def predict(self, u, i):
        with torch.no_grad():
            return forward(self.p, u, i, training=False, act=self.act, matmul=self.matmul)[0].numpy()
