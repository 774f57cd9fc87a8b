"""BASELINE.json configs[3]: NeuMF with row-sharded 20M x 2M x 64 tables, local batch 65536 per GPU.
Run under torchrun (one rank per GPU):  mode peer (NVLink peer loads / REDs inside the fused kernels) vs
mode nccl (explicit all-to-all of ids, rows and row gradients).  Device-timed, max over ranks."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from binrec_b200 import distributed as D
from binrec_b200.sharded import ShardedNeuMFNet

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    U = int(os.environ.get("SH_USERS", 20_000_000)); I = int(os.environ.get("SH_ITEMS", 2_000_000))
    E = 64; B = int(os.environ.get("SH_BATCH", 65536)); K = int(os.environ.get("SH_STEPS", 20))
    modes = os.environ.get("SH_MODES", "peer,nccl").split(",")
    res = {}
    for mode in modes:
        if mode == "nccl" and world == 1:
            continue
        for dist_name in ("uniform", "zipf"):
            net = ShardedNeuMFNet(U, I, E, dropout=0.2, device=dev, mode=mode, tensor_cores=os.environ.get("SH_TC", "1") == "1")
            g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
            nb = 8
            if dist_name == "uniform":
                us = torch.randint(0, U, (nb, B), generator=g, device=dev, dtype=torch.int32)
                its = torch.randint(0, I, (nb, B), generator=g, device=dev, dtype=torch.int32)
            else:   # heavy head: id = floor(N * r^3)
                us = (U * torch.rand((nb, B), generator=g, device=dev) ** 3).to(torch.int32)
                its = (I * torch.rand((nb, B), generator=g, device=dev) ** 3).to(torch.int32)
            y = (torch.rand((nb, B), generator=g, device=dev) < 0.2).float()
            o = torch.empty(B, device=dev); l = torch.empty(1, device=dev)
            def step(k):
                net.train_on_batch(us[k % nb], its[k % nb], y[k % nb], first_index=k * B * world, epoch=0, out=o, loss_out=l)
            for k in range(3):
                step(k)
            torch.cuda.synchronize()
            if world > 1: dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(K):
                step(3 + k)
            e1.record(); torch.cuda.synchronize()
            net.check()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item()) / K
            res[f"{mode}/{dist_name}"] = {"ms_per_step": ms, "interactions_per_s": world * B / (ms * 1e-3), "loss": float(l.item())}
            if rank == 0:
                print(mode, dist_name, res[f"{mode}/{dist_name}"], flush=True)
            del net, us, its, y
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({"config": f"NeuMF E=64 (MLP 128-64-32-16), U={U}, I={I}, row-sharded x{world}, local batch {B}, "
                                    f"lazy Adam, dropout 0.2", "n_gpus": world, "results": res}))
    if world > 1:
        dist.destroy_process_group()

main()
