"""One-launch NeuMF step (class graph F = 32, ML-1M tables, batch 16 384, Adam in the kernel) timed as bench.py's neumf block
does: L2 flushed before every step, CUDA events around the step -- and back to back."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.NeuMFModel import NeuMFNet
from binrec_b200 import hotpath as H
dev = torch.device("cuda:0")
U, I, B, K = 6040, 3706, 16384, 50
g = torch.Generator(device=dev); g.manual_seed(0)
for tag, kw in (("class graph F=32", dict(dropout=0.2, tensor_cores=True)),
                ("He et al. variant", dict(dropout=0.0, mf_dim=8, mf_mode="hadamard", batch_norm=False))):
    net = NeuMFNet(U, I, 32, device=dev, **kw)
    u = torch.randint(0, U, (B,), generator=g, device=dev, dtype=torch.int32)
    i = torch.randint(0, I, (B,), generator=g, device=dev, dtype=torch.int32)
    y = (torch.rand(B, generator=g, device=dev) < 0.2).float()
    step = lambda k: net.train_on_batch(u, i, y, first_index=k * B)
    for k in range(5):
        step(k)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev); sink = torch.empty((), dtype=torch.float32, device=dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for k in range(K):
        flush.zero_(); torch.sum(flush, dim=0, out=sink)
        ev[k][0].record(); step(5 + k); ev[k][1].record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        step(100 + k)
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: flushed L2 median {ts[K // 2]:.1f} us/step (min {ts[0]:.1f}); back to back {e0.elapsed_time(e1) * 1e3 / K:.1f} us/step")
