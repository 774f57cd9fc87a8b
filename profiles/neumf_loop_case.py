"""NeuMF fit inner loop at small batches (the reference trains with batch 128, NeuMFModel.py:102): per-step Python
calls (train_on_batch) against one C call for the whole list of batches (train_steps -> brk_neumf_train_steps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0")
U, I, n = 6040, 3706, 1 << 20
g = torch.Generator(device=dev); g.manual_seed(0)
u = torch.randint(0, U, (n,), generator=g, device=dev, dtype=torch.int32)
i = torch.randint(0, I, (n,), generator=g, device=dev, dtype=torch.int32)
y = (torch.rand(n, generator=g, device=dev) < 0.2).float()
for B in (128, 1024, 16384):
    for tcores in (False, True):
        net = NeuMFNet(U, I, 32, dropout=0.2, device=dev, tensor_cores=tcores)
        steps = min(400, n // B)
        order = np.arange(steps)
        out = torch.empty(B, device=dev); l = torch.empty(steps, device=dev)
        def loop():
            for k in range(steps):
                s = slice(k * B, (k + 1) * B)
                net.train_on_batch(u[s], i[s], y[s], first_index=k * B, out=out, loss_out=l[k:k + 1])
        def onecall():
            net.train_steps(u, i, y, B, order, losses=l, out=out)
        for name, fn in (("train_on_batch loop", loop), ("train_steps (one C call)", onecall)):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / steps
            print(f"B={B:6d} tensor_cores={int(tcores)} {name:26s}: {us:7.1f} us/step  {B / us:8.2f} M interactions/s", flush=True)
