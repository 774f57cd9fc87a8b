"""Diagnostic: device NeuMF gradients vs fp32 and fp64 autograd oracles (max abs error per parameter)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import neumf as ON
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0")
for (E, hidden, act, loss, dropout, B) in [(10, (100, 50, 10), "sigmoid", "bce", 0.0, 1000), (10, (100, 50, 10), "sigmoid", "bce", 0.2, 128),
                                           (16, (16, 8, 4), "sigmoid", "bce", 0.2, 128), (32, (32, 16, 8), "relu", "mse", 0.2, 1000)]:
    U, I = 300, 200
    o32 = ON.NeuMFOracle(U, I, emb=E, hidden=hidden, act=act, loss=loss, dropout=dropout)
    o64 = ON.NeuMFOracle(U, I, emb=E, hidden=hidden, act=act, loss=loss, dropout=dropout, dtype=torch.float64)
    net = NeuMFNet(U, I, E, hidden=hidden, act=act, loss=loss, dropout=dropout, device=dev)
    rng = np.random.default_rng(B + E)
    u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
    y = (rng.random(B) < 0.25).astype(np.float32)
    l32, out32, _ = o32.loss_and_grads(u, i, y, 4096, 3)
    l64, out64, _ = o64.loss_and_grads(u, i, y, 4096, 3)
    lg, og = net.forward_backward(*(torch.from_numpy(x).to(dev) for x in (u, i, y)), first_index=4096, epoch=3)
    print(f"== E={E} hidden={hidden} {act}/{loss} dropout={dropout} B={B}: loss dev {lg.item():.8f} o32 {float(l32):.8f} o64 {float(l64):.8f}")
    print("   out: dev-64 %.2e  o32-64 %.2e" % (np.abs(og.cpu().numpy() - out64.numpy()).max(), np.abs(out32.numpy() - out64.numpy()).max()))
    names = ["uMLP", "iMLP", "uMF", "iMF"] + list(net.DENSE_ORDER)
    for n in names:
        g64 = o64.p.t[n].grad.numpy(); g32 = o32.p.t[n].grad.numpy()
        gd = (net.tables()[names.index(n)].g if n in names[:4] else net.param(n, grad=True)).cpu().numpy().reshape(g64.shape)
        print(f"   {n:5s} |g|max {np.abs(g64).max():.2e}  dev-64 {np.abs(gd - g64).max():.2e}  o32-64 {np.abs(g32 - g64).max():.2e}")
