"""One launch each of the HBM-bound kernels at BASELINE.json configs[3] table sizes, for `ncu --set full`
(dram__bytes / gpu__time_duration -> measured HBM GB/s): gather_rows, scatter_add_rows, dense Keras Adam, lazy
Adam over touched rows, fused BPR fwd/bwd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200 import hotpath as H
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
g = torch.Generator(device=dev); g.manual_seed(0)
rows, d, n = 20_000_000, 64, 4_000_000
flush = torch.empty(256 << 18, dtype=torch.float32, device=dev)
table = torch.empty(rows, d, device=dev).uniform_(-1, 1)
ids = torch.randint(0, rows, (n,), generator=g, device=dev, dtype=torch.int32)
out = torch.empty(n, d, device=dev)
for _ in range(3):
    flush.zero_(); H.gather_rows(table, ids, out)
acc = torch.zeros(rows, d, device=dev)
touched = torch.zeros((rows + 31) // 32, dtype=torch.int32, device=dev)
for _ in range(3):
    flush.zero_(); H.scatter_add_rows(acc, ids, out, touched)
del out
tab = H.Table(table, g=acc); tab.touched = touched
opt = H.Adam(1e-3, sparse="lazy", device=dev)
for _ in range(2):
    H.scatter_add_rows(tab.g, ids[:1_000_000], torch.ones(1_000_000, d, device=dev), tab.touched)
    flush.zero_(); opt.apply([tab])                      # lazy Adam over ~1M touched rows
optd = H.Adam(1e-3, sparse="keras", device=dev)
small = H.Table(torch.empty(2_000_000, d, device=dev).uniform_(-1, 1), touched=False)
for _ in range(3):
    flush.zero_(); optd.apply([small])                   # dense Keras Adam over a 2M x 64 table
item = H.Table(torch.empty(2_000_000, d, device=dev).uniform_(-1, 1), touched=False)
B = 1_000_000
u = torch.randint(0, rows, (B,), generator=g, device=dev, dtype=torch.int32)
p = torch.randint(0, 2_000_000, (B,), generator=g, device=dev, dtype=torch.int32)
nn = torch.randint(0, 2_000_000, (B,), generator=g, device=dev, dtype=torch.int32)
tab.touched = None
for _ in range(3):
    flush.zero_(); H.bpr_fwd_bwd(tab, item, u, p, nn)
torch.cuda.synchronize()
print("done")
