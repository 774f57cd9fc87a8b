"""Where the full-catalog top-K kernel's epilogue stands against the TMEM read rate: the kernel as shipped, the same kernel
with the epilogue only draining TMEM (BRK_TOPK_PROBE=1: 32-bit reads), and draining it as packed 16-bit columns
(BRK_TOPK_PROBE=2: tcgen05.ld ... .pack::16b, 64 columns per instruction) -- the read an fp16-accumulator filter pass would do.
  python profiles/topk_probe.py [U I d]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200 import hotpath as H
dev = torch.device("cuda:0")
U, I, d = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 250_000, 64)
g = torch.Generator(device=dev); g.manual_seed(0)
Q = torch.randn(U, d, generator=g, device=dev); C = torch.randn(I, d, generator=g, device=dev)
idx = H.BruteForceIndex(10).index(C)
for mode in ("", "1", "2"):
    if mode:
        os.environ["BRK_TOPK_PROBE"] = mode
    else:
        os.environ.pop("BRK_TOPK_PROBE", None)
    for _ in range(2):
        idx(Q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        idx(Q)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"probe={mode or 0}: {ms:.3f} ms  {2.0 * U * I * max(64, (d + 63) // 64 * 64) / ms / 1e9:.0f} TFLOP/s  "
          f"{U * I / ms / 1e6 / 148 / 1.9:.1f} scores/clk/SM at 1.9 GHz")
# the fast path alone: all-positive queries and 16 leading items that beat everything -> thresholds saturate in the first tile
Qp = Q.abs(); Cs = C.clone(); Cs[:16] = 10.0
idx2 = H.BruteForceIndex(10).index(Cs)
for mode in ("",):
    os.environ.pop("BRK_TOPK_PROBE", None)
    for _ in range(2):
        idx2(Qp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        idx2(Qp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"saturated thresholds (no insertions after the first tile), probe={mode or 0}: {ms:.3f} ms  {U * I / ms / 1e6 / 148 / 1.9:.1f} scores/clk/SM")
