"""Short NeuMF case for ncu: 8 one-launch training steps (class graph numFactor 32, ML-1M tables, batch 16 384), then 4 steps
at BASELINE.json configs[3] table sizes (E = 64, 20 M x 2 M, batch 65 536: the five-kernel path + lazy Adam).
  ncu --set full -k regex:'fused_step|tc_|adam_rows' ... python profiles/neumf_ncu_case.py [small]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
U, I, B = 6040, 3706, 16384
net = NeuMFNet(U, I, 32, dropout=0.2, device=dev, tensor_cores=True)
u = torch.randint(0, U, (B,), generator=g, device=dev, dtype=torch.int32)
i = torch.randint(0, I, (B,), generator=g, device=dev, dtype=torch.int32)
y = (torch.rand(B, generator=g, device=dev) < 0.2).float()
for k in range(8):
    l, _ = net.train_on_batch(u, i, y, first_index=k * B)
torch.cuda.synchronize()
print("class graph loss", float(l.item()))
if len(sys.argv) > 1 and sys.argv[1] == "small":
    sys.exit(0)
del net
from binrec_b200.sharded import ShardedNeuMFNet
Uc, Ic, Bc = 20_000_000, 2_000_000, 65536
netc = ShardedNeuMFNet(Uc, Ic, 64, dropout=0.2, device=dev, mode="peer", tensor_cores=True)
us = torch.randint(0, Uc, (Bc,), generator=g, device=dev, dtype=torch.int32)
its = torch.randint(0, Ic, (Bc,), generator=g, device=dev, dtype=torch.int32)
yc = (torch.rand(Bc, generator=g, device=dev) < 0.2).float()
for k in range(4):
    l, _ = netc.train_on_batch(us, its, yc, first_index=k * Bc)
torch.cuda.synchronize()
print("configs[3] sizes loss", float(l.item()))
