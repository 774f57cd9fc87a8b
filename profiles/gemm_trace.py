"""Where a small gemm_tf32_kernel launch spends its time: %globaltimer stamps of CTA (0,0,0) at the phase boundaries
(brk_gemm_tf32_trace), for the products of the two-tower step at batch 1000, plus the launch's event-timed duration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from binrec_b200 import _native as Nn
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
NAMES = ["entry->tmem", "tmem->chunk0", "chunk0->mma done", "mma->acc in smem", "smem->stores", "stores->dealloc"]


def run(M, N, K, ta, tb, acc, label):
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev)
    C = torch.zeros(M, N, device=dev); tr = torch.zeros(8, dtype=torch.int64, device=dev)
    lda, ldb = A.shape[1], B.shape[1]
    lib, ctx = Nn.lib(), Nn.ctx(dev)
    traced = lambda: Nn.check(lib.brk_gemm_tf32_trace(ctx, Nn.ptr(A), Nn.ptr(B), Nn.ptr(C), None, M, N, K, lda, ldb, N, ta, tb, 1.0,
                                                      acc, Nn.ptr(tr), Nn.stream_ptr()), "trace")
    plain = lambda: Nn.check(lib.brk_gemm_tf32(ctx, Nn.ptr(A), Nn.ptr(B), Nn.ptr(C), None, M, N, K, lda, ldb, N, ta, tb, 1.0, acc,
                                               Nn.stream_ptr()), "gemm")
    for _ in range(5): plain()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): plain()
    e1.record(); torch.cuda.synchronize()
    rows = []
    for _ in range(5):
        traced(); torch.cuda.synchronize()
        t = tr.cpu().numpy()[:7]
        rows.append(np.diff(t))
    d = np.median(np.array(rows), axis=0)
    print(f"{label}: M={M} N={N} K={K} ta={ta} tb={tb} acc={acc}: {e0.elapsed_time(e1) * 1e3 / 50:.2f} us per back-to-back launch; "
          f"CTA 0 lifetime {d.sum() / 1e3:.2f} us: " + ", ".join(f"{n} {x / 1e3:.2f}" for n, x in zip(NAMES, d)), flush=True)


run(1000, 128, 128, 0, 0, 0, "tower Dense      ")
run(1000, 1000, 128, 0, 1, 0, "scores Q C^T     ")
run(1000, 128, 1000, 0, 0, 1, "dq = P C         ")
run(128, 128, 1000, 1, 0, 1, "dW = e^T dz      ")
run(1000, 128, 128, 0, 1, 0, "de = dz W^T      ")
