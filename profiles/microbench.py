"""Kernel micro-benchmarks (CUDA events, L2 flushed between iterations) used to fill the roofline
tables in DESIGN.md.  Run on the GPU box:  python profiles/microbench.py [section ...]
Sections: gather scatter bpr adam topk neumf."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from binrec_b200 import hotpath as H

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 18, dtype=torch.float32, device=dev)


def timeit(fn, iters=20, warm=3, do_flush=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e-3


def report(name, secs, nbytes):
    gbs = nbytes / secs / 1e9
    print(f"{name:64s} {secs*1e6:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK:6.3f} of measured HBM peak", flush=True)


def ids_of(kind, rows, n, g):
    if kind == "uniform":
        return torch.randint(0, rows, (n,), generator=g, device=dev, dtype=torch.int32)
    r = torch.rand(n, generator=g, device=dev)
    return (rows * r ** 3).to(torch.int32).clamp_(0, rows - 1)


def sec_gather():
    g = torch.Generator(device=dev); g.manual_seed(0)
    for rows, d, n in ((20_000_000, 64, 4_000_000), (2_000_000, 64, 4_000_000), (20_000_000, 32, 4_000_000),
                       (6040, 64, 1_000_000)):
        table = torch.empty(rows, d, device=dev).uniform_(-1, 1)
        out = torch.empty(n, d, device=dev)
        for kind in ("uniform", "skew"):
            ids = ids_of(kind, rows, n, g)
            s = timeit(lambda: H.gather_rows(table, ids, out))
            report(f"gather_rows rows={rows} d={d} n={n} ids={kind}", s, n * (2 * 4 * d + 4))
        del table, out


def sec_scatter():
    g = torch.Generator(device=dev); g.manual_seed(1)
    for rows, d, n in ((20_000_000, 64, 4_000_000), (2_000_000, 64, 4_000_000), (6040, 64, 1_000_000),
                       (6040, 64, 16384)):
        acc = torch.zeros(rows, d, device=dev)
        vals = torch.empty(n, d, device=dev).uniform_(-1, 1)
        for kind in ("uniform", "skew"):
            ids = ids_of(kind, rows, n, g)
            s = timeit(lambda: H.scatter_add_rows(acc, ids, vals))
            report(f"scatter_add_rows rows={rows} d={d} n={n} ids={kind}", s, n * (3 * 4 * d + 4))
            s = timeit(lambda: H.scatter_add_rows(acc, ids, vals, mode=1))
            report(f"scatter_add_rows mode=1 (sort+segment) rows={rows} d={d} n={n} ids={kind}", s, n * (3 * 4 * d + 4))
        del acc, vals


def sec_bpr():
    g = torch.Generator(device=dev); g.manual_seed(2)
    U, I, d = 6040, 3706, 64
    for B in (16384, 65536, 1_000_000):
        for touched in (True, False):
            for kind in ("uniform", "skew"):
                user = H.Table(torch.empty(U, d, device=dev).uniform_(-.05, .05), touched=touched)
                item = H.Table(torch.empty(I, d, device=dev).uniform_(-.05, .05), touched=touched)
                u, p, n = ids_of(kind, U, B, g), ids_of(kind, I, B, g), ids_of("uniform", I, B, g)
                loss = torch.empty(1, device=dev)
                s = timeit(lambda: H.bpr_fwd_bwd(user, item, u, p, n, loss))
                report(f"bpr_fwd_bwd ML-1M tables B={B} touched={touched} ids={kind}", s, B * 1536)
    U, I = 20_000_000, 2_000_000
    user = H.Table(torch.empty(U, d, device=dev).uniform_(-.05, .05), slots=0)
    item = H.Table(torch.empty(I, d, device=dev).uniform_(-.05, .05), slots=0)
    for B in (65536, 1_000_000):
        for kind in ("uniform", "skew"):
            u, p, n = ids_of(kind, U, B, g), ids_of(kind, I, B, g), ids_of("uniform", I, B, g)
            loss = torch.empty(1, device=dev)
            s = timeit(lambda: H.bpr_fwd_bwd(user, item, u, p, n, loss))
            report(f"bpr_fwd_bwd 20M x 2M tables B={B} ids={kind}", s, B * 1536)


def sec_adam():
    for rows, d in ((6040 + 3706, 64), (2_000_000, 64), (20_000_000, 64)):
        t = H.Table(torch.empty(rows, d, device=dev).uniform_(-.05, .05))
        opt = H.Adam(1e-3, device=dev)
        s = timeit(lambda: opt.apply([t]))
        report(f"adam_dense_keras rows={rows} d={d}", s, rows * d * 32)
        g = torch.Generator(device=dev); g.manual_seed(3)
        lazy = H.Adam(1e-3, sparse="lazy", device=dev)
        for n in (65536, 1_000_000):
            ids = torch.randint(0, rows, (n,), generator=g, device=dev, dtype=torch.int32)
            vals = torch.ones(n, d, device=dev)
            uniq = int(torch.unique(ids).numel())
            def run():
                H.scatter_add_rows(t.g, ids, vals, t.touched)
                lazy.apply([t])
            def only_scatter():
                H.scatter_add_rows(t.g, ids, vals, t.touched)
            s_both = timeit(run)
            t.g.zero_(); t.touched.zero_()
            s_sc = timeit(only_scatter)
            t.g.zero_(); t.touched.zero_()
            report(f"adam_rows rows={rows} d={d} touched={uniq} (scatter time subtracted)", max(s_both - s_sc, 1e-9),
                   uniq * d * 32 + rows // 8)
        del t


def sec_topk():
    """users/s of the fused scoring+top-K kernel; tensor-pipe fraction = 2*U*I*dpad / t / bf16 peak."""
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0}
    for U, I, d in ((6040, 3706, 128), (6040, 3706, 64), (131072, 250000, 64), (131072, 250000, 128),
                    (32768, 2000000, 64)):
        Q = torch.randn(U, d, device=dev); C = torch.randn(I, d, device=dev)
        idx = H.BruteForceIndex(10).index(C)
        q = H.rows_to_bf16(Q)
        # time the kernel alone (bf16 query conversion excluded, as the index is built once)
        from binrec_b200 import _native as N
        lib, ctx = N.lib(), N.ctx(dev)
        vals = torch.empty(U, 10, device=dev); ids = torch.empty(U, 10, dtype=torch.int32, device=dev)
        wsb = lib.brk_score_topk_workspace_bytes(ctx, U, I, 10)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        def run():
            N.check(lib.brk_score_topk_bf16(ctx, N.ptr(q), U, N.ptr(idx._c), I, q.shape[1], 10, 0, N.ptr(vals),
                                            N.ptr(ids), N.ptr(ws), wsb, N.stream_ptr()), "topk")
        s = timeit(run, iters=5, warm=2)
        flops = 2.0 * U * I * q.shape[1]
        print(f"score_topk U={U} I={I} d={d}: {s*1e3:9.3f} ms  {U/s/1e6:8.3f} M users/s  "
              f"{flops/s/1e12:7.1f} TFLOP/s = {flops/s/1e12/peaks['bf16_tflops']:.3f} of measured bf16 peak", flush=True)
        del Q, C, idx, q


def sec_neumf():
    """interactions/s of one NeuMF training step (5 fwd/bwd kernels + fused Adam), ML-1M tables."""
    from binrec_b200.NeuMFModel import NeuMFNet
    g = torch.Generator(device=dev); g.manual_seed(4)
    for E, B in ((32, 16384), (32, 65536), (64, 65536)):
        U, I = (6040, 3706) if E == 32 else (2_000_000, 200_000)
        net = NeuMFNet(U, I, E, dropout=0.2, device=dev)
        u, i = ids_of("skew", U, B, g), ids_of("skew", I, B, g)
        y = (torch.rand(B, generator=g, device=dev) < 0.2).float()
        out = torch.empty(B, device=dev); loss = torch.empty(1, device=dev)
        s = timeit(lambda: net.train_on_batch(u, i, y, out=out, loss_out=loss))
        s_fb = timeit(lambda: net.forward_backward(u, i, y, out=out, loss_out=loss))
        for t in net.tables() + [net.dense]:
            t.g.zero_()
        per = 4 * 4 * E * 2 + 16
        print(f"neumf_step E={E} B={B} tables {U}x{I}: step {s*1e6:8.1f} us ({B/s/1e6:7.1f} M inter/s)  fwd+bwd only "
              f"{s_fb*1e6:8.1f} us  -> {B*per/s_fb/1e9:7.1f} GB/s algorithmic = {B*per/s_fb/1e9/PEAK:.3f} of HBM peak; "
              f"{B*16182*(E/32)**2/s_fb/1e12:6.2f} TFLOP/s fp32", flush=True)
        del net


if __name__ == "__main__":
    secs = sys.argv[1:] or ["gather", "scatter", "bpr", "adam"]
    print(torch.cuda.get_device_name(0), "HBM peak (measured)", PEAK, "GB/s")
    for s in secs:
        globals()["sec_" + s]()
