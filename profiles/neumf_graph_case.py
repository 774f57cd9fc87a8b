"""NeuMF step replayed from a CUDA graph vs launched eagerly (small batch: launch-bound?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
U, I = 6040, 3706
for tcf in (False, True):
    net = NeuMFNet(U, I, E, dropout=0.2, device=dev, tensor_cores=tcf)
    g = torch.Generator(device=dev); g.manual_seed(0)
    u = torch.randint(0, U, (B,), generator=g, device=dev, dtype=torch.int32)
    i = torch.randint(0, I, (B,), generator=g, device=dev, dtype=torch.int32)
    y = (torch.rand(B, generator=g, device=dev) < 0.2).float()
    o = torch.empty(B, device=dev); l = torch.empty(1, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5): net.train_on_batch(u, i, y, out=o, loss_out=l)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): net.train_on_batch(u, i, y, out=o, loss_out=l)
        e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) * 1e3 / 50
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            net.train_on_batch(u, i, y, out=o, loss_out=l)
        for _ in range(5): gr.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50): gr.replay()
        e1.record(); torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) * 1e3 / 50
    print(f"E={E} B={B} tensor_cores={tcf}: eager {eager:.1f} us/step, graph {graph:.1f} us/step, loss {l.item():.5f}", flush=True)
