"""Why is bench.py's e2e copy leg slower than profiles/e2e_warm.py?  Replays bench.py's sequence with switches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from binrec_b200 import synth
from binrec_b200.BPRModel import BPRNet
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
B, K, W = 16384, 600, 5
users, items = synth.make_interactions(); U, I = synth.ML1M_USERS, synth.ML1M_ITEMS
nb = len(users) // B
net = BPRNet(U, I, 64, seed=42, device=dev); net.set_training_pairs(users, items); net.sample_negatives(7, 0)
order = [k % nb for k in range(K)]
hl = torch.empty(K + W, dtype=torch.float32).pin_memory()
packed = BPRNet.pack_host_batches(users[:nb * B], items[:nb * B], B)
def timed(tag):
    net.train_steps_from_host(packed, None, order[:W], B, 7, 1, hl[:W]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); net.train_steps_from_host(packed, None, order, B, 7, 1, hl[W:W + K]); e1.record()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{tag}: enqueue {1e6*(t1-t0)/K:.2f} us/step device {1e3*e0.elapsed_time(e1)/K:.2f} wall {1e6*(t2-t0)/K:.2f}", flush=True)
timed("fresh")
flush = torch.empty(bench.L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)
lb = torch.empty(1, device=dev)
for k in range(100):
    flush.zero_(); net.train_steps([k % nb], B, losses=lb)
torch.cuda.synchronize()
timed("after flushed steps")
net.train_steps(order, B); torch.cuda.synchronize()
timed("after hot loop")
s = bench.ClockSampler(0); s.start()
timed("with clock sampler"); print(s.stop())
hu = torch.from_numpy(users[:nb * B].copy()).pin_memory(); hp = torch.from_numpy(items[:nb * B].copy()).pin_memory()
timed("after more pinned allocations")
