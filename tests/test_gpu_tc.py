"""tcgen05 / TF32 building blocks of the NeuMF tensor-core path: every descriptor form used by
csrc/neumf_tc.cu (K-major and MN-major views of 128-byte-swizzled fp32 tiles) against a host product.
Inputs are multiples of 1/8 with few bits, so TF32 operands and fp32 accumulation are exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [(128, 64, 0), (128, 32, 0), (128, 16, 0), (128, 128, 0), (64, 32, 0),
         (128, 64, 3), (128, 32, 3), (64, 32, 3), (64, 16, 3), (128, 64, 1), (128, 64, 2)]


def _run(dev, M, N, K, mode, A, B):
    from binrec_b200 import _native as Nn
    out = torch.full((128, N), float("nan"), dtype=torch.float32, device=dev)
    Ad, Bd = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    Nn.check(Nn.lib().brk_tc_selftest(Nn.ctx(dev), M, N, K, mode, Nn.ptr(Ad), A.shape[0], A.shape[1], Nn.ptr(Bd), B.shape[0],
                                      B.shape[1], Nn.ptr(out), Nn.stream_ptr()), "brk_tc_selftest")
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("M,N,mode", CASES)
def test_tf32_descriptor_forms(dev, M, N, mode):
    rng = np.random.default_rng(M * 1000 + N * 10 + mode)
    K = 128
    a_mn, b_mn = mode & 1, (mode >> 1) & 1
    pad32 = lambda x: (x + 31) // 32 * 32
    Al = (rng.integers(-8, 9, size=(M, K)) / 8.0).astype(np.float32)      # logical A [M, K]
    Bl = (rng.integers(-8, 9, size=(N, K)) / 8.0).astype(np.float32)      # logical B [N, K]
    if a_mn:
        A = np.zeros((K, pad32(M)), np.float32); A[:, :M] = Al.T
    else:
        A = Al.copy()
    if b_mn:
        B = np.zeros((K, pad32(N)), np.float32); B[:, :N] = Bl.T
    else:
        B = np.zeros((max(N, 8), K), np.float32); B[:N] = Bl
    got = _run(dev, M, N, K, mode, A, B)
    ref = Al.astype(np.float64) @ Bl.astype(np.float64).T
    if M == 128:
        assert np.array_equal(got, ref.astype(np.float32))
    else:
        # M = 64: find where the 64 rows land among the 128 TMEM lanes, then require exactness
        lanes = [int(np.where((got == ref[m].astype(np.float32)).all(axis=1))[0][0]) for m in range(M)]
        print("M=64 row -> lane:", lanes[:8], "...", lanes[-4:])
        assert lanes == list(range(64)) or lanes == [32 * (m // 16) + m % 16 for m in range(64)]
