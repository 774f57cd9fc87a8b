"""One fused scoring+top-K call at a C5-shard-like size (for ncu).  python profiles/topk_case.py [U I d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200 import hotpath as H
U, I, d = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (32768, 250000, 64)
dev = torch.device("cuda:0")
Q = torch.randn(U, d, device=dev); C = torch.randn(I, d, device=dev)
idx = H.BruteForceIndex(10).index(C)
for _ in range(3):
    v, i = idx(Q)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); v, i = idx(Q); e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1), v[0, :3].tolist())
