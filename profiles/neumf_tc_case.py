"""NeuMF training step time, fp32 CUDA-core path vs TF32 tensor-core path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
U, I = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (6040, 3706)
lazy = U > 100000
for tcf in (False, True):
    net = NeuMFNet(U, I, E, dropout=0.2, device=dev, tensor_cores=tcf, sparse_adam="lazy" if lazy else "keras")
    g = torch.Generator(device=dev); g.manual_seed(0)
    u = torch.randint(0, U, (B,), generator=g, device=dev, dtype=torch.int32)
    i = torch.randint(0, I, (B,), generator=g, device=dev, dtype=torch.int32)
    y = (torch.rand(B, generator=g, device=dev) < 0.2).float()
    o = torch.empty(B, device=dev); l = torch.empty(1, device=dev)
    for _ in range(5):
        net.train_on_batch(u, i, y, out=o, loss_out=l)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        net.train_on_batch(u, i, y, out=o, loss_out=l)
    e1.record(); torch.cuda.synchronize()
    print(f"tensor_cores={tcf}: us/step {e0.elapsed_time(e1) * 1e3 / 20:.1f}  ({B / (e0.elapsed_time(e1) * 1e-3 / 20) / 1e6:.1f} M/s)  loss {l.item():.6f}", flush=True)
    del net
