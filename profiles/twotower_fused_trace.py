"""One-launch two-tower step (csrc/twotower_fused.cu): block-0 phase stamps (BRK_TT_TRACE) and the step time replayed as a
CUDA graph (what TwoTowerModel.fit does), beside the multi-kernel step (BRK_TT_NO_FUSED=1) at the same shape.
  python profiles/twotower_fused_trace.py [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.twoTower import TwoTowerModel
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
U, I = 6040, 3706
NAMES = ["entry", "towers (phase F)", "barrier", "score tile + row partials", "barrier", "lse, P, dq / dc products + REDs", "barrier",
         "dW / de products + REDs"]


def build():
    m = TwoTowerModel(128, I, U, "u", "i", list(range(U)), list(range(I)), semb=128, device=dev, tensor_cores=True)
    m.compile("Adagrad", learningRate=0.1)
    return m


def timed_graph(m, uid, iid, n=50):
    def step():
        m._train_ids(uid, iid, None)
    step()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


gen = torch.Generator(device=dev); gen.manual_seed(0)
uid = torch.randint(2, U + 2, (B,), generator=gen, device=dev, dtype=torch.int32)
iid = torch.randint(2, I + 2, (B,), generator=gen, device=dev, dtype=torch.int32)
trace = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ["BRK_TT_TRACE"] = hex(trace.data_ptr())
m = build()
for _ in range(3):
    l = m._train_ids(uid, iid, None)
torch.cuda.synchronize()
t = trace.cpu().numpy()
print(f"one-launch two-tower step, E = S = 128, batch {B}: loss {l.item():.4f}; block 0 phases (us since entry; delta)")
for k in range(1, 8):
    print(f"  {NAMES[k]:40s} {(t[k] - t[0]) / 1e3:8.2f}  (+{(t[k] - t[k - 1]) / 1e3:6.2f})")
print(f"  Adagrad on touched rows + Dense blocks (after a 4th barrier) {(t[15] - t[7]) / 1e3:8.2f}")
print(f"  inside the score phase: tiles staged {(t[8] - t[2]) / 1e3:.2f}, product {(t[9] - t[8]) / 1e3:.2f}, read-out + mask + row partials {(t[10] - t[9]) / 1e3:.2f}")
print(f"  inside the gradient phase: lse {(t[11] - t[4]) / 1e3:.2f}, P tile written {(t[12] - t[11]) / 1e3:.2f}, dq product {(t[13] - t[12]) / 1e3:.2f}, "
      f"P transposed view + dc product + dq REDs {(t[14] - t[13]) / 1e3:.2f}, dc REDs {(t[5] - t[14]) / 1e3:.2f}")
del os.environ["BRK_TT_TRACE"]
print(f"step + Adagrad as one graph replay: {timed_graph(build(), uid, iid):.1f} us")
os.environ["BRK_TT_NO_FUSED"] = "1"
print(f"multi-kernel step + Adagrad as one graph replay: {timed_graph(build(), uid, iid):.1f} us")
