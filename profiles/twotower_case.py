"""Two-tower training step (E = S = 128, in-batch softmax, Adagrad) time and launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200.twoTower import TwoTowerModel
dev = torch.device("cuda:0")
U, I = 6040, 3706
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
tc_flag = (sys.argv[2] == "tc") if len(sys.argv) > 2 else False
tt = TwoTowerModel(128, I, U, "u", "i", list(range(U)), list(range(I)), semb=128, device=dev, tensor_cores=tc_flag)
tt.compile("Adagrad", learningRate=0.1)
g = torch.Generator(device=dev); g.manual_seed(0)
uid = torch.randint(2, U + 2, (Bt,), generator=g, device=dev, dtype=torch.int32)
iid = torch.randint(2, I + 2, (Bt,), generator=g, device=dev, dtype=torch.int32)
def step():
    tt._step(uid, iid, None, True)
    tt.optimizer.apply([tt.userTower.emb, tt.itemTower.emb], dense=[tt.userTower.dense, tt.itemTower.dense])
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): step()
e1.record(); torch.cuda.synchronize()
print(f"tensor_cores={tc_flag} B={Bt}: {e0.elapsed_time(e1) * 1e3 / 50:.1f} us/step  {Bt / (e0.elapsed_time(e1) * 1e-3 / 50) / 1e6:.2f} M interactions/s")
# the same step captured once into a CUDA graph (the fork/join inside brk_twotower_step becomes two graph branches)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    step()
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    step()
for _ in range(5): graph.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(50): graph.replay()
e1.record(); torch.cuda.synchronize()
print(f"  as a CUDA graph: {e0.elapsed_time(e1) * 1e3 / 50:.1f} us/step  {Bt / (e0.elapsed_time(e1) * 1e-3 / 50) / 1e6:.2f} M interactions/s")
