"""ORACLE / CPU BASELINE (test infrastructure, never the product path): the BPR training step of
oracle/bpr.py restated with torch CPU tensor ops so that it can use all host cores.  This is the
"reference CPU loop" bench.py times (cpu_baseline.kind = "port"): the reference's own loop is
Keras/TensorFlow (/root/reference/src/models/BPRModel.py:109), which cannot be installed here --
this is NOT TensorFlow.  Same math, same Keras-Adam (dense-equivalent) semantics, fp32.
tests/test_oracle_models.py checks it against the NumPy oracle.
"""
import numpy as np
import torch


class BPRTorchCPU:
    def __init__(self, user, item, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        self.user = torch.from_numpy(np.array(user, dtype=np.float32, copy=True))
        self.item = torch.from_numpy(np.array(item, dtype=np.float32, copy=True))
        self.mu = torch.zeros_like(self.user); self.vu = torch.zeros_like(self.user)
        self.mi = torch.zeros_like(self.item); self.vi = torch.zeros_like(self.item)
        self.gu = torch.zeros_like(self.user); self.gi = torch.zeros_like(self.item)
        self.t = 0
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps

    def _adam(self, w, m, v, g, alpha):
        m.mul_(self.b1).add_(g, alpha=1.0 - self.b1)
        v.mul_(self.b2).addcmul_(g, g, value=1.0 - self.b2)
        w.addcdiv_(m, v.sqrt().add_(self.eps), value=-alpha)

    @torch.no_grad()
    def step(self, u, p, n):
        u = torch.as_tensor(u, dtype=torch.int64); p = torch.as_tensor(p, dtype=torch.int64)
        n = torch.as_tensor(n, dtype=torch.int64)
        B = u.numel()
        ue, pe, ne = self.user.index_select(0, u), self.item.index_select(0, p), self.item.index_select(0, n)
        diff = pe - ne
        x = (ue * diff).sum(-1)
        s = torch.sigmoid(x)
        loss = (1.0 - s).mean()
        g = (-(s * (1.0 - s)) / B).unsqueeze(1)
        self.gu.zero_(); self.gi.zero_()
        self.gu.index_add_(0, u, g * diff)
        gue = g * ue
        self.gi.index_add_(0, p, gue)
        self.gi.index_add_(0, n, -gue)
        self.t += 1
        alpha = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        self._adam(self.user, self.mu, self.vu, self.gu, alpha)
        self._adam(self.item, self.mi, self.vi, self.gi, alpha)
        return float(loss)
