"""Back-to-back BPR steps through the multi-step C entry point (hot L2), for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from binrec_b200 import synth
from binrec_b200.BPRModel import BPRNet
dev = torch.device("cuda:0")
users, items = synth.make_interactions()
net = BPRNet(6040, 3706, 64, device=dev)
net.set_training_pairs(users, items); net.sample_negatives(7, 0)
order = list(range(40))
net.train_steps(order, 16384); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); net.train_steps(order, 16384); e1.record(); torch.cuda.synchronize()
print("us/step", e0.elapsed_time(e1) * 1e3 / len(order))
