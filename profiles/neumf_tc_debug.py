import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0")
U, I = 300, 200
for E in (64, 32):
  for act, loss in (("sigmoid", "bce"), ("relu", "mse")):
    for dropout in (0.0, 0.2):
      for B in (1000, 128, 77):
        ref = NeuMFNet(U, I, E, act=act, loss=loss, dropout=dropout, device=dev, tensor_cores=False)
        net = NeuMFNet(U, I, E, act=act, loss=loss, dropout=dropout, device=dev, tensor_cores=True)
        rng = np.random.default_rng(E + B)
        u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
        y = (rng.random(B) < 0.25).astype(np.float32)
        ud, idd, yd = (torch.from_numpy(x).to(dev) for x in (u, i, y))
        l0, o0 = ref.forward_backward(ud, idd, yd, first_index=4096, epoch=3)
        l1, o1 = net.forward_backward(ud, idd, yd, first_index=4096, epoch=3)
        worst_rel, worst_frac, worst_max = 0, 1, 0
        wn = ""
        items = [(n, getattr(ref, n).g.cpu().numpy(), getattr(net, n).g.cpu().numpy()) for n in ("uMLP", "iMLP", "uMF", "iMF")]
        items += [(n, ref.param(n, grad=True).cpu().numpy(), net.param(n, grad=True).cpu().numpy()) for n in net.DENSE_ORDER]
        for n, g0, g1 in items:
            scale = max(np.abs(g0).max(), 1e-20)
            rel = np.linalg.norm((g1 - g0).ravel()) / max(np.linalg.norm(g0.ravel()), 1e-20)
            mx = np.abs(g1 - g0).max() / scale
            if rel > worst_rel: worst_rel, wn = rel, n
            worst_max = max(worst_max, mx)
        print(f"E={E} {act:7s} drop={dropout} B={B:4d}: out err {float((o1-o0).abs().max()):.2e} loss rel {abs(l1.item()-l0.item())/abs(l0.item()):.2e} "
              f"worst frob-rel {worst_rel:.3f} ({wn}) worst max-err/max|g| {worst_max:.3f}")
