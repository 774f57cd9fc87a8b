"""ORACLE (test infrastructure, never the product path): BPR matrix factorisation restated on
the CPU.

Pinning.  WIRING PINNED: the reference's BPRModel.compileModel, bprTripletLoss and identityLoss are EXECUTED over a
torch-backed Keras stand-in (tests/golden/keras_shim.py, tests/golden/make_wiring_golden.py -> wiring_golden.npz) and
this oracle reproduces the per-triplet output, the loss, both table gradients (one item table shared by the positive
and the negative lookup) and the tables after the compiled Adam(1e-3) step to 1e-9 (tests/test_oracle_wiring.py).
UPSTREAM NUMERICS UNPINNED: Embedding / Adam arithmetic is restated from the upstream documentation (no TensorFlow
here, no golden vectors in the reference), held to hand-computed fp64 cases
(tests/test_oracle_models.py).

Follows /root/reference/src/models/BPRModel.py:49-74 (graph: one user table, one item table
shared by the positive and negative lookup), :124-144 (bprTripletLoss, identityLoss) and the
script twin /root/reference/src/models/bpr.py:136-192.

  x_b   = sum_d u_b*p_b - sum_d u_b*n_b              (BPRModel.py:139-142)
  l_b   = 1 - sigmoid(x_b)                           (BPRModel.py:144; NOT -log sigmoid)
  loss  = mean_b l_b                                 (identityLoss, BPRModel.py:126)
  dl/dx = -s(1-s)/B;  du = g (p-n), dp = g u, dn = -g u
"""
import numpy as np

from . import embedding as E


def bpr_forward(user_tab, item_tab, u, p, n):
    ue, pe, ne = user_tab[u], item_tab[p], item_tab[n]
    x = (ue * pe).sum(-1) - (ue * ne).sum(-1)
    s = 1.0 / (1.0 + np.exp(-x.astype(np.float64)))
    return x, s.astype(user_tab.dtype)


def bpr_loss_and_grads(user_tab, item_tab, u, p, n):
    """Returns (loss, dense grad of user table, dense grad of item table)."""
    dt = user_tab.dtype
    u = np.asarray(u, dtype=np.int64); p = np.asarray(p, dtype=np.int64); n = np.asarray(n, dtype=np.int64)
    B = len(u)
    ue, pe, ne = user_tab[u], item_tab[p], item_tab[n]
    x, s = bpr_forward(user_tab, item_tab, u, p, n)
    loss = dt.type(np.mean(1.0 - s.astype(np.float64)))
    g = (-(s * (1 - s)) / dt.type(B)).astype(dt)[:, None]
    gu = E.scatter_add_rows(user_tab.shape[0], u, g * (pe - ne), dtype=dt)
    gi = E.scatter_add_rows(item_tab.shape[0], p, g * ue, dtype=dt)
    np.add.at(gi, n, -g * ue)
    return loss, gu, gi


class BPROracle:
    """Training loop state: tables + Adam moments.  optimizer 'adam_keras' is what the reference
    runs (dense-equivalent sparse Adam); 'adam_lazy' is the row-sparse variant."""

    def __init__(self, num_users, num_items, dim, seed=42, dtype=np.float32, lr=1e-3,
                 optimizer="adam_keras"):
        rng = np.random.Generator(np.random.Philox(key=seed))
        self.user = E.keras_embedding_init(rng, num_users, dim, dtype)
        self.item = E.keras_embedding_init(rng, num_items, dim, dtype)
        self.mu = np.zeros_like(self.user); self.vu = np.zeros_like(self.user)
        self.mi = np.zeros_like(self.item); self.vi = np.zeros_like(self.item)
        self.t = 0
        self.lr = lr
        self.optimizer = optimizer

    def step(self, u, p, n):
        loss, gu, gi = bpr_loss_and_grads(self.user, self.item, u, p, n)
        self.t += 1
        if self.optimizer == "adam_keras":
            E.adam_dense_keras(self.user, self.mu, self.vu, gu, self.t, lr=self.lr)
            E.adam_dense_keras(self.item, self.mi, self.vi, gi, self.t, lr=self.lr)
        elif self.optimizer == "adam_lazy":
            E.adam_rows_lazy(self.user, self.mu, self.vu, gu, np.unique(u), self.t, lr=self.lr)
            E.adam_rows_lazy(self.item, self.mi, self.vi, gi,
                             np.unique(np.concatenate([p, n])), self.t, lr=self.lr)
        else:
            raise ValueError(self.optimizer)
        return loss


def bpr_predict(user_tab, item_tab, user_id, item_ids):
    """bpr.py:122-133: user vector times item matrix."""
    return item_tab[np.asarray(item_ids, dtype=np.int64)] @ user_tab[user_id]


def auc_and_ap_at_k(scores, positives, actual_len, k):
    """Per-user AUC and average precision at k as /root/reference/src/models/bpr.py:230-289 computes them, from one
    row of scores over the catalog: AUC = sklearn's roc_auc_score (ties count half); AP@k over the catalog sorted by
    score, descending, stable (earlier position first among equals), divided by min(actual_len, k).
    positives: distinct column positions.  Returns (auc, ap); NaN where the reference would raise."""
    s = np.asarray(scores)
    pos = np.asarray(sorted(set(int(p) for p in positives)), dtype=np.int64)
    is_pos = np.zeros(len(s), dtype=bool); is_pos[pos] = True
    neg = s[~is_pos]
    if len(pos) and len(neg):
        auc = float(np.mean([(np.sum(neg < s[p]) + 0.5 * np.sum(neg == s[p])) / len(neg) for p in pos]))
    else:
        auc = float("nan")
    order = np.argsort(-s.astype(np.float64), kind="stable")[:k]
    hits, score = 0.0, 0.0
    for rank, j in enumerate(order):
        if is_pos[j]:
            hits += 1.0
            score += hits / (rank + 1.0)
    den = min(int(actual_len), int(k))
    return auc, (score / den if den > 0 else float("nan"))
