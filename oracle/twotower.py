"""ORACLE (test infrastructure, never the product path): the two-tower model restated on the CPU with
torch autograd.

Pinning.  WIRING PINNED: the reference's TwoTowerModel (constructor, computeEmb, computeLossTfrs / computeLossRdZero,
train_step with its GradientTape / apply_gradients sequence, setCandidates + call) is EXECUTED over a torch-backed
Keras / TFRS stand-in (tests/golden/keras_shim.py, tests/golden/make_wiring_golden.py -> wiring_golden.npz) and this
oracle reproduces the loss, every gradient, the weights after the Adagrad(0.1) step and the BruteForce top-k lists to
1e-9 (tests/test_oracle_wiring.py; both loss modes; StringLookup offset 2; candidate_ids = the MATERIAL column).
UPSTREAM NUMERICS UNPINNED: what tfrs.tasks.Retrieval, BruteForce, Adagrad and the Keras layers compute inside is
restated from the upstream documentation (tensorflow_recommenders is not even listed in requirements.txt, SURVEY.md
section 8c).

Follows /root/reference/trainers/twoTower.py:19-111:
  towers : StringLookup -> Embedding(n + 2, E) -> Dense(S), linear (:33-41).  StringLookup (TF 2.3/2.4)
           maps vocabulary entry j to index j + 2 (0 = mask, 1 = out-of-vocabulary).
  loss   : tfrs.tasks.Retrieval(loss=None) called with candidate_ids (:47,82-83): scores = Q C^T,
           labels = identity, accidental hits get finfo(float32).min/100 added, categorical
           cross-entropy from logits, reduction SUM.
           rdZero (:85-87): sigmoid(Dot(q, c)) against RATING_TYPE, Keras BinaryCrossentropy (mean).
  update : Keras Adagrad(0.1) (:278-279), initial accumulator 0.1, eps 1e-7.
"""
import numpy as np
import torch

MIN_FLOAT = float(np.finfo(np.float32).min) / 100.0


class TwoTowerOracle:
    def __init__(self, n_users, n_items, E, S, seed=42, lr=0.1, dtype=torch.float32, rdZero=False, matmul=torch.matmul):
        # matmul: oracle/tf32.matmul gives the TF32-operand form of every product the tensor-core step computes
        # (tower Dense layers and the in-batch score matrix, forward and both gradient products each)
        self.mm = matmul
        rng = np.random.Generator(np.random.Philox(key=seed))
        npdt = np.float64 if dtype == torch.float64 else np.float32

        def emb(rows):
            return rng.uniform(-0.05, 0.05, size=(rows, E)).astype(np.float32)

        def glorot(i, o):
            lim = np.sqrt(6.0 / (i + o))
            return rng.uniform(-lim, lim, size=(i, o)).astype(np.float32)

        arrs = dict(Eu=emb(n_users + 2), Ei=emb(n_items + 2), Wu=glorot(E, S), bu=np.zeros(S, np.float32),
                    Wi=glorot(E, S), bi=np.zeros(S, np.float32))
        self.t = {k: torch.tensor(v.astype(npdt), requires_grad=True) for k, v in arrs.items()}
        self.acc = {k: torch.full_like(v, 0.1) for k, v in self.t.items()}
        self.lr, self.dtype, self.rdZero = lr, dtype, rdZero

    def towers(self, u_idx, i_idx):
        t = self.t
        q = self.mm(t["Eu"][torch.as_tensor(u_idx, dtype=torch.int64)], t["Wu"]) + t["bu"]
        c = self.mm(t["Ei"][torch.as_tensor(i_idx, dtype=torch.int64)], t["Wi"]) + t["bi"]
        return q, c

    def loss(self, u_idx, i_idx, cand_ids=None, labels=None):
        q, c = self.towers(u_idx, i_idx)
        if self.rdZero:
            logit = (q * c).sum(1)
            y = torch.as_tensor(np.asarray(labels), dtype=self.dtype)
            return torch.nn.functional.binary_cross_entropy_with_logits(logit, y)
        scores = self.mm(q, c.T)
        B = scores.shape[0]
        if cand_ids is not None:
            ids = torch.as_tensor(np.asarray(cand_ids))
            dup = (ids[:, None] == ids[None, :]).to(self.dtype) - torch.eye(B, dtype=self.dtype)
            scores = scores + dup * MIN_FLOAT
        return torch.nn.functional.cross_entropy(scores, torch.arange(B), reduction="sum")

    def loss_and_grads(self, u_idx, i_idx, cand_ids=None, labels=None):
        for v in self.t.values():
            v.grad = None
        l = self.loss(u_idx, i_idx, cand_ids, labels)
        l.backward()
        return l.detach()

    def step(self, u_idx, i_idx, cand_ids=None, labels=None):
        l = self.loss_and_grads(u_idx, i_idx, cand_ids, labels)
        with torch.no_grad():
            for k, w in self.t.items():
                g = w.grad if w.grad is not None else torch.zeros_like(w)
                self.acc[k] += g * g
                w -= self.lr * g / (self.acc[k].sqrt() + 1e-7)
        return float(l)

    def user_vectors(self, u_idx):
        with torch.no_grad():
            return (self.t["Eu"][torch.as_tensor(u_idx, dtype=torch.int64)] @ self.t["Wu"] + self.t["bu"]).numpy()

    def item_vectors(self, i_idx):
        with torch.no_grad():
            return (self.t["Ei"][torch.as_tensor(i_idx, dtype=torch.int64)] @ self.t["Wi"] + self.t["bi"]).numpy()
