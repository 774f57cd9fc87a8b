"""The headline e2e at the driver's --steps 20: BPRNet.train_steps_from_host / train_steps_mapped over 20 steps of 16 384 --
enqueue time (the C call returns), device time (CUDA events) and wall time to the synchronize, ten repetitions."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from binrec_b200 import synth
from binrec_b200.BPRModel import BPRNet
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
B, W = 16384, 5
users, items = synth.make_interactions(); U, I = synth.ML1M_USERS, synth.ML1M_ITEMS
nb = len(users) // B
net = BPRNet(U, I, 64, seed=42, device=dev); net.set_training_pairs(users, items); net.sample_negatives(7, 0)
packed = BPRNet.pack_host_batches(users[:nb * B], items[:nb * B], B)
hu = torch.from_numpy(users[:nb * B].copy()).pin_memory(); hp = torch.from_numpy(items[:nb * B].copy()).pin_memory()
for K in (20, 600):
    order = [k % nb for k in range(K)]
    hl = torch.empty(K + W, dtype=torch.float32).pin_memory()
    for tag, fn in (("copy  ", lambda o, l: net.train_steps_from_host(packed, None, o, B, 7, 1, l)),
                    ("mapped", lambda o, l: net.train_steps_mapped(hu, hp, o, B, 7, 1, l))):
        fn(order[:W], hl[:W]); torch.cuda.synchronize()
        rows = []
        for rep in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record(); fn(order, hl[W:W + K]); e1.record()
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            rows.append((1e6 * (t1 - t0), 1e3 * e0.elapsed_time(e1), 1e6 * (t2 - t0)))
        print(f"   walls in call order (us): {[round(r[2]) for r in rows]}")
        rows.sort(key=lambda r: r[2])
        enq, devt, wall = rows[len(rows) // 2]
        print(f"K={K:3d} {tag}: enqueue {enq:7.1f} us  device {devt:7.1f} us  wall {wall:7.1f} us  -> {K * B / wall:.0f} M interactions/s "
              f"({wall / K:.2f} us/step)", flush=True)
# the same 20-step copy call as isolated single shots (13 ms of idle in front of each, as in bench.py where one shot is timed),
# without and with bench.py's NVML clock sampler thread (a sample every 50 ms)
import bench
K = 20; order = [k % nb for k in range(K)]; hl = torch.empty(K + W, dtype=torch.float32).pin_memory()
for tag in ("no sampler", "NVML sampler"):
    smp = None
    if tag != "no sampler":
        smp = bench.ClockSampler(0); smp.start(); time.sleep(0.2)
    rows = []
    for rep in range(60):
        torch.cuda.synchronize(); time.sleep(0.013)
        t0 = time.perf_counter(); net.train_steps_from_host(packed, None, order, B, 7, 1, hl[W:W + K]); torch.cuda.synchronize(); t2 = time.perf_counter()
        rows.append(round(1e6 * (t2 - t0)))
    srt = sorted(rows)
    print(f"isolated shots, {tag}: median {srt[30]} us, p10 {srt[6]}, p90 {srt[54]}, first five {rows[:5]}", smp.stop() if smp else "")
