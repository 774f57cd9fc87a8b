"""Phase timeline of the fused data-parallel BPR kernel (globaltimer stamps of block 0), run under torchrun."""
import os, sys, ctypes
os.environ["BRK_COOP_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist, io, contextlib
from binrec_b200 import synth
from binrec_b200.BPRModel import BPRNet
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
users, items = synth.make_interactions()
net = BPRNet(synth.ML1M_USERS, synth.ML1M_ITEMS, 64, device=dev)
net.set_training_pairs(users, items); net.sample_negatives(7, 0)
B, K = 16384, 200
order = [(k * world + rank) % (len(users) // B) for k in range(K)]
net.train_steps(order[:8], B); torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); net.train_steps(order, B); e1.record(); torch.cuda.synchronize()
print(f"rank {rank}: {1e3 * e0.elapsed_time(e1) / K:.2f} us/step", flush=True)
from binrec_b200 import _native as N
buf = (ctypes.c_uint64 * (K * 8))()
lib = N.lib(); lib.brk_coop_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int32]
assert lib.brk_coop_trace_read(buf, K * 8) == 0
a = np.frombuffer(buf, dtype=np.uint64).reshape(K, 8).astype(np.int64)[20:]
names = ["phase1 (fwd/bwd + block 0 collects the local arrivals)", "barrier A (flags over NVLink)", "reduce + Adam of my slice", "block 0 collects the local arrivals", "barrier B (flags over NVLink)", "zero g + pull + loss + grid barrier"]
d = np.diff(a[:, :7], axis=1)
for r in range(world):
    if r == rank:
        print(f"rank {rank} mean ns per phase:", {n: int(x) for n, x in zip(names, d.mean(axis=0))}, "step", int((a[1:, 0] - a[:-1, 0]).mean()), flush=True)
    dist.barrier()
dist.destroy_process_group()
