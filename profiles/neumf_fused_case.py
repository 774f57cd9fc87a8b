"""NeuMF training step time: one-launch cooperative tensor-core kernel (csrc/neumf_fused.cu) vs the five-kernel
tensor-core path (csrc/neumf_tc.cu) vs fp32, step = forward/backward + exact Keras Adam over all parameters.
Usage: neumf_fused_case.py [E B U I]; also times forward/backward alone and the He et al. variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
U, I = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (6040, 3706)
lazy = U > 100000
N_IT = 50


def timeit(fn, n=N_IT):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def case(tag, env, **kw):
    if env:
        os.environ["BRK_NEUMF_NO_FUSED"] = "1"
    else:
        os.environ.pop("BRK_NEUMF_NO_FUSED", None)
    net = NeuMFNet(U, I, E, device=dev, sparse_adam="lazy" if lazy else "keras", **kw)
    g = torch.Generator(device=dev); g.manual_seed(0)
    u = torch.randint(0, U, (B,), generator=g, device=dev, dtype=torch.int32)
    i = torch.randint(0, I, (B,), generator=g, device=dev, dtype=torch.int32)
    y = (torch.rand(B, generator=g, device=dev) < 0.2).float()
    o = torch.empty(B, device=dev); l = torch.empty(1, device=dev)
    t_step = timeit(lambda: net.train_on_batch(u, i, y, out=o, loss_out=l))
    t_fb = timeit(lambda: net.forward_backward(u, i, y, out=o, loss_out=l))
    net.grad_arena.zero_()
    print(f"{tag:34s} E={E} B={B}: step {t_step:7.1f} us ({B / t_step:7.1f} M/s)   fwd/bwd alone {t_fb:7.1f} us   loss {l.item():.6f}",
          flush=True)


case("fp32 (neumf2.cu)", False, dropout=0.2, tensor_cores=False)
case("tensor cores, five kernels", True, dropout=0.2, tensor_cores=True)
case("tensor cores, one launch (fused)", False, dropout=0.2, tensor_cores=True)
if E == 32:
    case("He et al. variant (fused)", False, dropout=0.0, mf_dim=8, mf_mode="hadamard", batch_norm=False)
