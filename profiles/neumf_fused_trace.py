"""Phase stamps (%globaltimer, block 0) of the one-launch NeuMF step (csrc/neumf_fused.cu, BRK_NEUMF_TRACE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.NeuMFModel import NeuMFNet
dev = torch.device("cuda:0")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
U, I = 6040, 3706
NAMES = ["entry", "setup done", "x0 gathered+staged", "MMA1 done", "h1 + BN1 stats (grid.sync)", "h2 + BN2 stats (grid.sync)",
         "head done", "da2/dW3 + BN2-bwd sums (grid.sync)", "da1/dW2 + BN1-bwd sums (grid.sync)", "dx0 REDs issued",
         "slots stored, barrier, dense grads summed", "Adam over tables + exit"]
for tag, kw in (("class spec", dict(dropout=0.2, tensor_cores=True)),
                ("He variant", dict(dropout=0.0, mf_dim=8, mf_mode="hadamard", batch_norm=False))):
    if E != 32 and tag == "He variant":
        continue
    trace = torch.zeros(32 + 480, dtype=torch.int64, device=dev)
    os.environ["BRK_NEUMF_TRACE"] = hex(trace.data_ptr())
    net = NeuMFNet(U, I, E, device=dev, **kw)
    g = torch.Generator(device=dev); g.manual_seed(0)
    u = torch.randint(0, U, (B,), generator=g, device=dev, dtype=torch.int32)
    i = torch.randint(0, I, (B,), generator=g, device=dev, dtype=torch.int32)
    y = (torch.rand(B, generator=g, device=dev) < 0.2).float()
    for _ in range(5):
        net.train_on_batch(u, i, y)
    torch.cuda.synchronize()
    t = trace.cpu().numpy()
    print(f"{tag} E={E} B={B}: block 0 phases (us since entry; delta)")
    for k in range(1, 12):
        print(f"  {NAMES[k]:44s} {(t[k] - t[0]) / 1e3:8.2f}  (+{(t[k] - t[k - 1]) / 1e3:6.2f})")
    sm = t[32:32 + 148]
    print(f"  CTAs 0..147 run on {len(set(sm.tolist()))} distinct SMs")
    if t[12]:
        print(f"  inside the BN2 barrier phase: h2 read {(t[12] - t[4]) / 1e3:.2f}, tile sums {(t[13] - t[12]) / 1e3:.2f}, "
              f"grid barrier {(t[14] - t[13]) / 1e3:.2f}, slot reduction {(t[15] - t[14]) / 1e3:.2f}, stats {(t[5] - t[15]) / 1e3:.2f}")
    del os.environ["BRK_NEUMF_TRACE"]
