import os, sys
os.environ["BRK_GEMM_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from binrec_b200 import _native as Nn
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
M = N = K = 128
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.zeros(M, N, device=dev)
f = lambda: Nn.check(Nn.lib().brk_gemm_tf32(Nn.ctx(dev), Nn.ptr(A), Nn.ptr(B), Nn.ptr(C), None, M, N, K, K, K, N, 0, 1, 1.0, 0, Nn.stream_ptr()), "gemm")
for _ in range(5): f()
os.environ["BRK_GEMM_TRACE_PRINT"] = "1"
for _ in range(4): f()
