import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from binrec_b200 import _native as Nn
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
np.set_printoptions(linewidth=250, threshold=100000)
def run(M, N, K, mode, A, B):
    out = torch.full((128, N), float("nan"), dtype=torch.float32, device=dev)
    Ad, Bd = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    Nn.check(Nn.lib().brk_tc_selftest(Nn.ctx(dev), M, N, K, mode, Nn.ptr(Ad), A.shape[0], A.shape[1], Nn.ptr(Bd), B.shape[0],
                                      B.shape[1], Nn.ptr(out), Nn.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    return out.cpu().numpy()
M, N, K = 128, 64, 128
B = np.zeros((N, K), np.float32)
for n in range(N): B[n, n] = 1.0            # D[m][n] = A(m, k=n)
Ak = np.tile(np.arange(K, dtype=np.float32)[:, None], (1, M))     # tile [K rows][M cols], value = k
Am = np.tile(np.arange(M, dtype=np.float32)[None, :], (K, 1))     # value = m
dk = run(M, N, K, 1, Ak, B); dm = run(M, N, K, 1, Am, B)
print("value=k: expect D[m][n]=n"); print(dk[:10, :40].astype(int)); print(dk[60:68, :40].astype(int))
print("value=m: expect D[m][n]=m"); print(dm[:10, :40].astype(int)); print(dm[28:40, :12].astype(int)); print(dm[120:, :12].astype(int))
