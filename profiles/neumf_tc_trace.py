"""Phase stamps (%globaltimer, block 0 / thread 0) of the five-kernel NeuMF step at BASELINE.json configs[3] table sizes
(E = 64, 20 M x 2 M rows, batch 65 536, one GPU): csrc/neumf_tc.cu, BRK_NTC_TRACE.  Prints per-tile phase times of
tc_head and tc_bwd1 for the first tiles block 0 owns."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
dev = torch.device("cuda:0")
trace = torch.zeros(320, dtype=torch.int64, device=dev)
os.environ["BRK_NTC_TRACE"] = hex(trace.data_ptr())
from binrec_b200.sharded import ShardedNeuMFNet
g = torch.Generator(device=dev); g.manual_seed(0)
Uc, Ic, Bc = 20_000_000, 2_000_000, 65536
if len(sys.argv) > 1 and sys.argv[1] == "small":
    Uc, Ic = 6040, 3706
netc = ShardedNeuMFNet(Uc, Ic, 64, dropout=0.2, device=dev, mode="peer", tensor_cores=True)
us = torch.randint(0, Uc, (Bc,), generator=g, device=dev, dtype=torch.int32)
its = torch.randint(0, Ic, (Bc,), generator=g, device=dev, dtype=torch.int32)
yc = (torch.rand(Bc, generator=g, device=dev) < 0.2).float()
for k in range(4):
    l, _ = netc.train_on_batch(us, its, yc, first_index=k * Bc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(20):
    netc.train_on_batch(us, its, yc, first_index=(4 + k) * Bc)
e1.record(); torch.cuda.synchronize()
print(f"step (fwd/bwd + lazy Adam, back to back): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
t = trace.cpu().numpy()


def show(name, base, labels):
    """labels: what ends at each per-tile stamp after the tile-top stamp"""
    w = t[base:base + 64]
    n = int((w != 0).sum())
    print(f"{name}: {n} stamps, block 0 alive {(w[n - 1] - w[0]) / 1e3:.2f} us; set-up {(w[1] - w[0]) / 1e3:.2f}")
    per = len(labels) + 1
    k, tile = 1, 0
    while k + per <= n:
        parts = "  ".join(f"{nm} {(w[k + j + 1] - w[k + j]) / 1e3:5.2f}" for j, nm in enumerate(labels))
        nxt = (w[k + per] - w[k + per - 1]) / 1e3 if k + per < n else 0.0
        print(f"  tile {tile}: {parts}  end-of-tile sync {nxt:5.2f} | {(w[k + per] - w[k]) / 1e3 if k + per < n else 0:6.2f}")
        k += per
        tile += 1
    for j in range(k, n - 1):
        print(f"  tail +{(w[j + 1] - w[j]) / 1e3:.2f}")


show("tc_head", 0, ["ids", "stage d2", "MMA + MF gather", "logits", "dW3/dW4/dd2", "MF REDs + sync"])
show("tc_bwd1", 128, ["ids+masks", "gather x0", "stage dz", "MMAs", "dx0 REDs"])
