import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200 import _native as Nn
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
def run(M, N, K, ta, tb, iters=50):
    A = torch.randn((K, M) if ta else (M, K), device=dev); B = torch.randn((N, K) if tb else (K, N), device=dev)
    C = torch.zeros(M, N, device=dev)
    f = lambda: Nn.check(Nn.lib().brk_gemm_tf32(Nn.ctx(dev), Nn.ptr(A), Nn.ptr(B), Nn.ptr(C), None, M, N, K, A.shape[1], B.shape[1], N,
                                                ta, tb, 1.0, 0, Nn.stream_ptr()), "gemm")
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"M={M} N={N} K={K} ta={ta} tb={tb}: {us:.1f} us  {2.0*M*N*K/us/1e6:.1f} TFLOP/s", flush=True)
run(128, 128, 128, 0, 1); run(1000, 128, 128, 0, 0); run(1000, 1000, 128, 0, 1); run(8192, 8192, 128, 0, 1); run(8192, 128, 8192, 0, 0)
