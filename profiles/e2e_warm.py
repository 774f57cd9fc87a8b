"""How the host-fed path warms up: 5-step call, then repeated 600-step calls (as bench.py's e2e leg does)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from binrec_b200 import synth
from binrec_b200.BPRModel import BPRNet
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
B, K = 16384, 600
users, items = synth.make_interactions()
U, I = synth.ML1M_USERS, synth.ML1M_ITEMS
nb = len(users) // B
net = BPRNet(U, I, 64, seed=42, device=dev)
net.set_training_pairs(users, items); net.sample_negatives(7, 0)
hu = torch.from_numpy(users[:nb * B].copy()).pin_memory(); hp = torch.from_numpy(items[:nb * B].copy()).pin_memory()
order = [k % nb for k in range(K)]
hl = torch.empty(K + 5, dtype=torch.float32).pin_memory()
def run(name, fn, o):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record(); fn(o); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    k = len(o)
    print(f"{name:10s} K={k:4d} enqueue {1e6*(t1-t0)/k:7.2f} us/step   device {1e3*e0.elapsed_time(e1)/k:7.2f} us/step   "
          f"wall {1e6*(t2-t0)/k:7.2f} us/step", flush=True)
which = sys.argv[1] if len(sys.argv) > 1 else "copy"
packed = BPRNet.pack_host_batches(users[:nb * B], items[:nb * B], B)
f = (lambda o: net.train_steps_from_host(hu, hp, o, B, 7, 1, hl[:len(o)])) if which == "copy" else \
    (lambda o: net.train_steps_from_host(packed, None, o, B, 7, 1, hl[:len(o)])) if which == "packed" else \
    (lambda o: net.train_steps_mapped(hu, hp, o, B, 7, 1, hl[:len(o)]))
net.train_steps(order[:5], B)
run(which, f, order[:5])
for _ in range(4):
    run(which, f, order)
run(which, f, order[:61])
run(which, f, order[:61])
