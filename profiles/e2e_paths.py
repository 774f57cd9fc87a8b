"""Host-fed BPR training (bench.py's e2e leg) three ways, with the host enqueue cost separated from the
device time: per-step explicit copies (train_steps_from_host), zero-copy single launch
(train_steps_mapped), and the device-resident loop for reference."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from binrec_b200 import synth
from binrec_b200.BPRModel import BPRNet
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
B, K = 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 600
users, items = synth.make_interactions()
U, I = synth.ML1M_USERS, synth.ML1M_ITEMS
nb = len(users) // B
net = BPRNet(U, I, 64, seed=42, device=dev)
net.set_training_pairs(users, items); net.sample_negatives(7, 0)
hu = torch.from_numpy(users[:nb * B].copy()).pin_memory(); hp = torch.from_numpy(items[:nb * B].copy()).pin_memory()
order = [k % nb for k in range(K)]
hl = torch.empty(K, dtype=torch.float32).pin_memory()
def run(name, fn):
    fn(order[:8]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(order); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:10s} enqueue {1e6*(t1-t0)/K:7.2f} us/step   device {1e3*e0.elapsed_time(e1)/K:7.2f} us/step   "
          f"wall {1e6*(t2-t0)/K:7.2f} us/step  -> {K*B/(t2-t0)/1e6:8.1f} M/s", flush=True)
for rep in range(2):
    run("resident", lambda o: net.train_steps(o, B))
    run("copy", lambda o: net.train_steps_from_host(hu, hp, o, B, 7, 1, hl[:len(o)]))
    run("mapped", lambda o: net.train_steps_mapped(hu, hp, o, B, 7, 1, hl[:len(o)]))
print("losses finite:", bool(np.isfinite(hl.numpy()).all()), hl[:3].tolist())
# where the extra time of the fused-sampler kernel goes: device ids / host ids x device losses / host losses
du, dp = hu.to(dev), hp.to(dev); dl = torch.empty(K, dtype=torch.float32, device=dev)
run("smp dev/dev", lambda o: net.train_steps_mapped(du, dp, o, B, 7, 1, dl[:len(o)]))
run("smp dev/hst", lambda o: net.train_steps_mapped(du, dp, o, B, 7, 1, hl[:len(o)]))
run("smp hst/dev", lambda o: net.train_steps_mapped(hu, hp, o, B, 7, 1, dl[:len(o)]))
