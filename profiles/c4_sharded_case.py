"""BASELINE.json configs[3] row-sharded NeuMF step alone (what bench.py's c4_neumf_sharded block times), for quick A/B runs:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 profiles/c4_sharded_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
peaks = {"hbm_gbs": 6455.6}
out = bench.c4_sharded_block(dev, world, rank, peaks)
if rank == 0:
    print(f"configs[3] row-sharded x{world}: {out['ms_per_step']:.3f} ms/step  {out['value'] / 1e6:.1f} M interactions/s  loss {out['loss']:.5f}")
dist.barrier()
dist.destroy_process_group()
