"""BASELINE.json configs[4]: full-catalog top-K sweep, users x 2M items, items range-sharded over the ranks
(torchrun).  Every rank scores ALL users against its item range with the tcgen05 scoring kernel (fused top-K
epilogue), the per-shard [U, k] lists are all-gathered and merged (score desc, id asc).  Device-timed, max over ranks."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from binrec_b200 import hotpath as H, distributed as D

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    U = int(os.environ.get("TK_USERS", 1_000_000)); I = int(os.environ.get("TK_ITEMS", 2_000_000))
    k = 10; chunk = 131072
    res = {}
    for d in (64, 128):
        g = torch.Generator(device=dev); g.manual_seed(5)                # same queries on every rank
        Q = torch.randn(U, d, generator=g, device=dev)
        lo, hi = D.local_slice(I)
        gi = torch.Generator(device=dev); gi.manual_seed(100 + rank)
        C = torch.randn(hi - lo, d, generator=gi, device=dev)
        idx = H.BruteForceIndex(k).index(C, id_offset=lo)
        def sweep():
            outs = []
            for s in range(0, U, chunk):
                v, i = idx(Q[s:s + chunk])
                pv, pi = D.gather_topk_parts(v, i)
                outs.append(H.topk_merge(pv, pi))
            return outs
        sweep(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); outs = sweep(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dpad = (d + 63) // 64 * 64
        res[f"d{d}"] = {"ms": ms, "users_per_s": U / (ms * 1e-3), "tflops_per_gpu": 2.0 * U * (hi - lo) * dpad / (ms * 1e-3) / 1e12}
        if rank == 0: print(d, res[f"d{d}"], flush=True)
        del Q, C, idx, outs
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"config": f"top-K sweep {U} users x {I} items, k=10, bf16 tcgen05 scoring, items range-sharded x{world}, "
                                    f"merge after all-gather of [U,k] lists", "n_gpus": world, "results": res}))
    if world > 1: dist.destroy_process_group()
main()
