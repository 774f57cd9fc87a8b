"""Event-timed rates of the input-pipeline kernels (csrc/pipeline.cu): the ML-1M epoch frame (5 M rows), a
100 M-row frame, the epoch permutation alone, and the id factorisation at 20 M keys.  L2 flushed between
iterations.  Algorithmic bytes: epoch build 12 B written + 8 B of source pair read per row; factorisation
8 B key read + 4 B id written per key."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from binrec_b200 import pipeline as PL, synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
flush = torch.empty(256 << 18, dtype=torch.float32, device=dev)

def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

pu, pi = synth.make_interactions()
P = len(pu)
indptr, sitems = synth.build_csr(pu, pi, synth.ML1M_USERS)
t = lambda a: torch.from_numpy(a).to(dev)
dpu, dpi, dip, dsi = t(pu), t(pi), t(indptr), t(sitems)
n = 5 * P
out = (torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
       torch.empty(n, dtype=torch.float32, device=dev))
for reject in (False, True):
    ms = timed(lambda: PL.neumf_epoch_build(dpu, dpi, 4 * P, 7, 0, reject=reject, csr_indptr=dip, csr_items=dsi, out=out))
    print(f"neumf_epoch_build ML-1M rows={n} reject={reject}: {ms*1e3:.1f} us  {n/ms/1e6:.2f} G rows/s  {20*n/ms/1e6:.1f} GB/s algorithmic")
t0 = time.time()
import pandas as pd
df = pd.DataFrame({"u": pu, "i": pi})
neg = df.sample(frac=4., replace=True).copy(); neg.i = neg.i.sample(frac=1.).values
df["label"] = 1.; neg["label"] = 0.
merged = pd.concat([df, neg]).sample(frac=1.)
print(f"pandas bootstrapDataset frame (NeuMFModel.py:103-109) on the host: {(time.time()-t0)*1e3:.0f} ms")

Pb = 20_000_000
g = torch.Generator(device=dev); g.manual_seed(0)
bu = torch.randint(0, 20_000_000, (Pb,), generator=g, device=dev, dtype=torch.int32)
bi = torch.randint(0, 2_000_000, (Pb,), generator=g, device=dev, dtype=torch.int32)
nb = 5 * Pb
outb = (torch.empty(nb, dtype=torch.int32, device=dev), torch.empty(nb, dtype=torch.int32, device=dev),
        torch.empty(nb, dtype=torch.float32, device=dev))
ms = timed(lambda: PL.neumf_epoch_build(bu, bi, 4 * Pb, 7, 0, out=outb), 3)
print(f"neumf_epoch_build rows={nb} (20 M positives, 20 M x 2 M ids): {ms:.2f} ms  {nb/ms/1e6:.2f} G rows/s  {20*nb/ms/1e6:.1f} GB/s algorithmic")
perm = torch.empty(nb, dtype=torch.int64, device=dev)
ms = timed(lambda: PL.epoch_permutation(nb, 7, 0, 0, device=dev, out=perm), 3)
print(f"epoch_permutation n={nb}: {ms:.2f} ms  {nb/ms/1e6:.2f} G/s")
del outb, perm

for nk, distinct in ((20_000_000, 3_000_000), (20_000_000, 20_000_000), (1_000_209, 6040)):
    raw = torch.randint(0, distinct, (nk,), generator=g, device=dev)
    keys = raw * 2654435761 + 12345
    v = PL.Vocabulary(dev)
    ms = timed(lambda: v.build(keys), 3)
    t0 = time.time(); pd.factorize(keys.cpu().numpy()); host = time.time() - t0
    print(f"Vocabulary.build n={nk} distinct~{distinct}: {ms:.2f} ms (incl. allocations + one host read)  {nk/ms/1e6:.2f} G keys/s  "
          f"{12*nk/ms/1e6:.1f} GB/s algorithmic; pandas.factorize on the host: {host*1e3:.0f} ms")
    ms = timed(lambda: v.lookup(keys), 3)
    print(f"Vocabulary.lookup n={nk}: {ms:.2f} ms  {nk/ms/1e6:.2f} G keys/s")
print("done")
