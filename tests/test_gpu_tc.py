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


# ---- which TF32 does the tensor core compute? ----------------------------------------------------------------------
TF32_MODE = "trunc"      # oracle/tf32.py: the 13 dropped mantissa bits are truncated (determined by the test below)


def test_tf32_operand_rounding_is_truncation(dev):
    """A has full fp32 mantissas, B is exactly representable: the product equals trunc(A) @ B and differs from
    round-to-nearest(A) @ B -- this pins the operand rule oracle/tf32.py restates."""
    from oracle import tf32
    rng = np.random.default_rng(11)
    M, N, K = 128, 64, 128
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.integers(-8, 9, size=(N, K)) / 8.0).astype(np.float32)
    got = _run(dev, M, N, K, 0, A, B).astype(np.float64)
    ref_t = tf32.trunc_np(A).astype(np.float64) @ B.astype(np.float64).T
    ref_r = tf32.rn_np(A).astype(np.float64) @ B.astype(np.float64).T
    err_t, err_r = np.abs(got - ref_t).max(), np.abs(got - ref_r).max()
    print("TF32 operand rule: max |err| vs truncation", err_t, "vs round-to-nearest", err_r)
    assert (err_t < 2e-5) != (err_r < 2e-5), (err_t, err_r)
    assert (err_t < err_r) == (TF32_MODE == "trunc"), (err_t, err_r)


# ---- NeuMF on the tensor cores vs the TF32-operand ORACLE ----------------------------------------------------------
# oracle/neumf.py with matmul = oracle/tf32.py: every Dense product (forward, input gradient, weight gradient) reads
# its operands with 10 mantissa bits and accumulates in fp32 -- the arithmetic of tcgen05.mma.kind::tf32.  Against
# that oracle the device is held to (SURVEY.md section 8c / VERDICT r1 item 2a):
#   predictions rtol 2e-3 / atol 1e-5, loss rtol 2e-3, BatchNorm moving statistics rtol 2e-3;
#   gradients per tensor: max |error| <= 1e-2 of the tensor's largest entry and Frobenius error <= 1e-2 (embedding
#   tables: up to 3 rows may exceed the max-norm bound, by at most 0.1 -- see the gate-flip note below; it happens once,
#   at batch 16 384: one row of 6040).
# One more case is kept apart (test_neumf_relu_gate_flip_case_is_bounded): ReLU is discontinuous and TF32 truncation
# turns a 1-ulp difference of an activation (summation order) into a 2^-10 step of the operand, so a pre-activation that
# lands within that noise of zero flips its gate in one of the two implementations and changes that ONE sample's whole
# contribution.  The batch seeded 1064 (E = 64, B = 1000, dropout) contains such a sample: W1 is off by 1.2e-2
# Frobenius there, the rows of that sample by up to 5e-2, in BOTH device paths alike; that case is held to 5e-2 / 0.1.
# The measured errors are printed; typical: predictions ~1e-6, gradients ~1e-4 of the largest entry.
# path: "fused" = the one-launch cooperative kernel (csrc/neumf_fused.cu), "five" = the five-kernel path
# (csrc/neumf_tc.cu, BRK_NEUMF_NO_FUSED), which is what runs when a batch does not fit on chip.
def _tf32_matmul():
    from oracle import tf32
    return tf32.matmul if TF32_MODE == "trunc" else tf32.matmul_rn


def _grad_err(g1, g0, floor=0.0):
    """(max |err| / max |g|, Frobenius error, rows whose max |err| exceeds 1e-2 max |g|, rows).  floor: lower bound of the
    scale -- a bias that feeds a BatchNorm has a true gradient of ZERO (the layer subtracts the batch mean), what both
    implementations hold there is rounding noise, so such tensors are judged against the scale of the whole gradient."""
    scale = max(float(np.abs(g0).max()), floor, 1e-20)
    e = np.abs(g1 - g0).reshape(g0.shape[0], -1) if g0.ndim > 1 else np.abs(g1 - g0).reshape(1, -1)
    bad = int((e.max(axis=1) > 1e-2 * scale).sum())
    return (float(e.max()) / scale,
            float(np.linalg.norm((g1 - g0).ravel()) / max(np.linalg.norm(g0.ravel()), floor * np.sqrt(g0.size), 1e-20)),
            bad, e.shape[0])


def _path(monkeypatch, path):
    if path == "five":
        monkeypatch.setenv("BRK_NEUMF_NO_FUSED", "1")
    else:
        monkeypatch.delenv("BRK_NEUMF_NO_FUSED", raising=False)


def _check_step_against_oracle(dev, net, orc, u, i, y, first, epoch, tag, tol=(1e-2, 1e-2)):
    lref, oref, aux = orc.loss_and_grads(u, i, y, first_index=first, epoch=epoch)
    ud, idd, yd = (torch.from_numpy(x).to(dev) for x in (u, i, y))
    lgot, ogot = net.forward_backward(ud, idd, yd, first_index=first, epoch=epoch)
    np.testing.assert_allclose(ogot.cpu().numpy(), oref.numpy(), rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(lgot.item(), float(lref), rtol=2e-3)
    worst = (0.0, 0.0, "", 0)
    names = ["uMLP", "iMLP", "uMF", "iMF"] + list(net.DENSE_ORDER)
    gmax = max(float(orc.p.t[n].grad.abs().max()) for n in net.DENSE_ORDER if orc.p.t[n].grad is not None)
    for name in names:
        g0 = orc.p.t[name].grad
        g0 = np.zeros(tuple(orc.p.t[name].shape), np.float32) if g0 is None else g0.numpy()
        g1 = (getattr(net, name).g if name in ("uMLP", "iMLP", "uMF", "iMF") else net.param(name, grad=True)).cpu().numpy()
        g1 = g1.reshape(g0.shape)
        if not np.abs(g0).max() > 0:                      # unused BN slots of the variant: gradient exactly zero
            assert not np.abs(g1).max() > 0, name
            continue
        mx, rel, bad, rows = _grad_err(g1, g0, floor=1e-2 * gmax if (name in ("b1", "b2") and net.batch_norm) else 0.0)
        worst = max(worst, (mx, rel, name, bad))
        # embedding tables: a gate flip of ONE sample shows as a few rows (its user / item) -- at most 3 rows may exceed
        # the max-norm bound, and never by more than 0.1 of the largest entry
        row_budget = name in ("uMLP", "iMLP", "uMF", "iMF") and bad <= 3 and mx <= 0.1
        assert rel <= tol[0] and (mx <= tol[1] or row_budget), (tag, name, mx, rel, bad, rows)
    print(f"{tag}: pred max|err| {np.abs(ogot.cpu().numpy() - oref.numpy()).max():.2e}, worst gradient {worst}")


@pytest.mark.parametrize("path", ["fused", "five"])
@pytest.mark.parametrize("E", [64, 32])
@pytest.mark.parametrize("dropout", [0.0, 0.2])
@pytest.mark.parametrize("B", [1000, 128, 77])
def test_neumf_tensor_core_step_matches_tf32_oracle(dev, monkeypatch, path, E, dropout, B):
    from binrec_b200.NeuMFModel import NeuMFNet
    from oracle import neumf as ON
    _path(monkeypatch, path)
    U, I = 300, 200
    net = NeuMFNet(U, I, E, dropout=dropout, seed=42, dropout_seed=11, device=dev, tensor_cores=True)
    orc = ON.NeuMFOracle(U, I, emb=E, seed=42, dropout=dropout, dropout_seed=11, matmul=_tf32_matmul())
    rng = np.random.default_rng(E + B + (2 if dropout else 0))
    u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
    y = (rng.random(B) < 0.25).astype(np.float32)
    _check_step_against_oracle(dev, net, orc, u, i, y, 4096, 3, f"{path} E={E} B={B} dropout={dropout}")
    assert not net._bufs["acc"].any().item()              # accumulators left zero for the next step


@pytest.mark.parametrize("path", ["fused", "five"])
def test_neumf_relu_gate_flip_case_is_bounded(dev, monkeypatch, path):
    """The batch with a ReLU gate inside TF32 truncation noise (see the comment above): bounded, not tight."""
    from binrec_b200.NeuMFModel import NeuMFNet
    from oracle import neumf as ON
    _path(monkeypatch, path)
    U, I, E, B = 300, 200, 64, 1000
    net = NeuMFNet(U, I, E, dropout=0.2, seed=42, dropout_seed=11, device=dev, tensor_cores=True)
    orc = ON.NeuMFOracle(U, I, emb=E, seed=42, dropout=0.2, dropout_seed=11, matmul=_tf32_matmul())
    rng = np.random.default_rng(E + B)
    u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
    y = (rng.random(B) < 0.25).astype(np.float32)
    _check_step_against_oracle(dev, net, orc, u, i, y, 4096, 3, f"{path} gate-flip case", tol=(5e-2, 0.1))


@pytest.mark.parametrize("dropout", [0.0, 0.2])
def test_neumf_five_kernel_sigmoid_bce_matches_tf32_oracle(dev, dropout):
    """The sigmoid / BCE graph (trainers/NFC_plain.py activations) exists on the five-kernel tensor-core path only."""
    from binrec_b200.NeuMFModel import NeuMFNet
    from oracle import neumf as ON
    U, I, B, E = 300, 200, 1000, 32
    net = NeuMFNet(U, I, E, act="sigmoid", loss="bce", dropout=dropout, device=dev, tensor_cores=True)
    orc = ON.NeuMFOracle(U, I, emb=E, act="sigmoid", loss="bce", dropout=dropout, matmul=_tf32_matmul())
    rng = np.random.default_rng(1)
    u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
    y = (rng.random(B) < 0.25).astype(np.float32)
    _check_step_against_oracle(dev, net, orc, u, i, y, 0, 1, f"five sigmoid/bce dropout={dropout}")


@pytest.mark.parametrize("B", [1000, 128, 16384])
def test_neumf_he_variant_matches_oracles(dev, B):
    """BASELINE.json configs[0]: GMF (8-dim Hadamard vector) + MLP 64-32-16-8, no BatchNorm, no dropout."""
    from binrec_b200.NeuMFModel import NeuMFNet
    from oracle import neumf as ON
    U, I = 6040, 3706
    kw = dict(mf_dim=8, mf_mode="hadamard", batch_norm=False)
    net = NeuMFNet(U, I, 32, dropout=0.0, device=dev, **kw)
    orc = ON.NeuMFOracle(U, I, emb=32, dropout=0.0, matmul=_tf32_matmul(), **kw)
    o32 = ON.NeuMFOracle(U, I, emb=32, dropout=0.0, **kw)
    assert np.array_equal(net.param("W4").cpu().numpy().reshape(-1), orc.p.t["W4"].detach().numpy().reshape(-1))
    assert net.param("W4").numel() == 8 + 8
    rng = np.random.default_rng(B)
    u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
    y = (rng.random(B) < 0.2).astype(np.float32)
    _check_step_against_oracle(dev, net, orc, u, i, y, 0, 0, f"He variant B={B}")
    l32, o32p, _ = o32.loss_and_grads(u, i, y)                # and the plain fp32 oracle, TF32 tolerance
    out, _ = net.predict_on_batch(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev))
    np.testing.assert_allclose(out.cpu().numpy(), o32p.numpy(), rtol=2e-3, atol=1e-4)


def test_neumf_baseline_shape_three_steps_match_tf32_oracle(dev):
    """BASELINE.json configs[0] shape: 6040 x 3706, numFactor 32, batch 16 384, three Keras-Adam steps with dropout."""
    from binrec_b200.NeuMFModel import NeuMFNet
    from oracle import neumf as ON
    U, I, B, E = 6040, 3706, 16384, 32
    net = NeuMFNet(U, I, E, dropout=0.2, device=dev, tensor_cores=True)
    orc = ON.NeuMFOracle(U, I, emb=E, dropout=0.2, matmul=_tf32_matmul())
    o64 = ON.NeuMFOracle(U, I, emb=E, dropout=0.2, dtype=torch.float64)
    rng = np.random.default_rng(2)
    for step in range(3):
        u = (U * rng.random(B) ** 1.5).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
        y = (rng.random(B) < 0.2).astype(np.float32)
        lref, _ = orc.step(u, i, y, first_index=step * B, epoch=0)
        l64, _ = o64.step(u, i, y, first_index=step * B, epoch=0)
        lgot, _ = net.train_on_batch(*(torch.from_numpy(x).to(dev) for x in (u, i, y)), first_index=step * B, epoch=0)
        np.testing.assert_allclose(lgot.item(), lref, rtol=2e-3)
        np.testing.assert_allclose(lgot.item(), l64, rtol=2e-3)
    ref, ref64 = orc.p.numpy(), o64.p.numpy()
    for name, tab in zip(("uMLP", "iMLP", "uMF", "iMF"), net.tables()):
        w = tab.w.cpu().numpy()
        # three Adam steps move a weight by at most 3 lr = 3e-3; the device must stay as close to the fp64 run as
        # the TF32 oracle does (x3 + 1e-4: Adam normalises the step, so gradient noise near |g| ~ eps is amplified)
        err_dev, err_orc = np.abs(w - ref64[name]).max(), np.abs(ref[name] - ref64[name]).max()
        print(name, "max |w - fp64|: device", err_dev, "tf32 oracle", err_orc)
        assert err_dev <= 3 * err_orc + 1e-4, (name, err_dev, err_orc)
    np.testing.assert_allclose(net.bn_moving.cpu().numpy(),
                               np.concatenate([orc.p.mm1.numpy(), orc.p.mv1.numpy(), orc.p.mm2.numpy(), orc.p.mv2.numpy()]),
                               rtol=2e-3, atol=1e-5)


def test_neumf_tensor_core_eval_uses_moving_statistics(dev):
    from binrec_b200.NeuMFModel import NeuMFNet
    from oracle import neumf as ON
    U, I, B = 300, 200, 2048
    net = NeuMFNet(U, I, 64, dropout=0.2, device=dev, tensor_cores=True)
    orc = ON.NeuMFOracle(U, I, emb=64, dropout=0.2, matmul=_tf32_matmul())
    rng = np.random.default_rng(9)
    for step in range(5):
        u = (U * rng.random(B) ** 2).astype(np.int32); i = (I * rng.random(B) ** 2).astype(np.int32)
        y = (rng.random(B) < 0.25).astype(np.float32)
        lref, _ = orc.step(u, i, y, first_index=step * B, epoch=0)
        lgot, _ = net.train_on_batch(*(torch.from_numpy(x).to(dev) for x in (u, i, y)), first_index=step * B, epoch=0)
        np.testing.assert_allclose(lgot.item(), lref, rtol=2e-3)
    p1, _ = net.predict_on_batch(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev), torch.from_numpy(y).to(dev))
    np.testing.assert_allclose(p1.cpu().numpy(), orc.predict(u, i), rtol=5e-3, atol=1e-4)


# ---- general TF32 product (csrc/gemm_tc.cu) and the two-tower step on it -------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (1000, 128, 128), (1000, 1000, 128), (1000, 128, 1000), (128, 128, 1000),
                                   (300, 64, 96), (77, 200, 40), (4100, 128, 64)])   # tile widths 32 / 64 / 128 all occur
@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_tf32_all_layouts_exact_on_representable_inputs(dev, M, N, K, ta, tb):
    from binrec_b200 import _native as Nn
    rng = np.random.default_rng(M + N + K + 2 * ta + tb)
    pad4 = lambda x: (x + 3) // 4 * 4
    Al = (rng.integers(-8, 9, size=(M, K)) / 8.0).astype(np.float32)
    Bl = (rng.integers(-8, 9, size=(K, N)) / 8.0).astype(np.float32)
    bias = (rng.integers(-8, 9, size=N) / 4.0).astype(np.float32)
    # storage with padded leading dimensions (multiples of 4 floats)
    if ta: As = np.zeros((K, pad4(M)), np.float32); As[:, :M] = Al.T
    else:  As = np.zeros((M, pad4(K)), np.float32); As[:, :K] = Al
    if tb: Bs = np.zeros((N, pad4(K)), np.float32); Bs[:, :K] = Bl.T
    else:  Bs = np.zeros((K, pad4(N)), np.float32); Bs[:, :N] = Bl
    Ad, Bd, bd = (torch.from_numpy(x).to(dev) for x in (As, Bs, bias))
    C0 = (rng.integers(-4, 5, size=(M, N)) / 2.0).astype(np.float32)
    for accumulate in (0, 1):
        Cd = torch.from_numpy(C0.copy()).to(dev)
        Nn.check(Nn.lib().brk_gemm_tf32(Nn.ctx(dev), Nn.ptr(Ad), Nn.ptr(Bd), Nn.ptr(Cd), Nn.ptr(bd) if not accumulate else None,
                                        M, N, K, As.shape[1], Bs.shape[1], N, ta, tb, 0.5, accumulate, Nn.stream_ptr()),
                 "brk_gemm_tf32")
        ref = 0.5 * (Al.astype(np.float64) @ Bl.astype(np.float64))
        ref = ref + (C0 if accumulate else bias[None, :])
        assert np.array_equal(Cd.cpu().numpy(), ref.astype(np.float32)), (accumulate,)


@pytest.mark.parametrize("U,I,E,S,B", [(500, 300, 128, 128, 1000), (6040, 3706, 128, 128, 1000), (300, 200, 64, 64, 2048),
                                       (300, 200, 100, 60, 333), (300, 200, 32, 128, 100), (300, 200, 128, 128, 1536),
                                       (300, 200, 64, 64, 129)])
def test_twotower_tensor_core_step_matches_tf32_oracle(dev, U, I, E, S, B):
    """The two-tower step with every product on tcgen05 (TF32 operands) against oracle/twotower.py with oracle/tf32.py's
    matmul: loss rtol 2e-3; gradients max |error| <= 1e-2 of the largest entry and Frobenius <= 1e-2 (no ReLU here: the graph
    is smooth, measured ~1e-4).  (6040, 3706, 128, 128, 1000) is BASELINE.json configs[2].  Batches up to 1536 run as the ONE
    cooperative launch of csrc/twotower_fused.cu (ragged last tiles, widths that are not multiples of 32, T = 1 / 2 / 3 / 8 /
    12 row blocks); batch 2048 takes the multi-kernel step of csrc/twotower.cu + gemm_tc.cu."""
    from binrec_b200.twoTower import TwoTowerModel
    from oracle import twotower as OT
    m = TwoTowerModel(E, I, U, "u", "i", list(range(U)), list(range(I)), semb=S, device=dev, tensor_cores=True)
    o = OT.TwoTowerOracle(U, I, E, S, seed=42, matmul=_tf32_matmul())
    with torch.no_grad():                                          # identical weights: copy the device model's into the oracle
        o.t["Eu"].copy_(m.userTower.emb.w.cpu()); o.t["Ei"].copy_(m.itemTower.emb.w.cpu())
        for tw, wn, bn in ((m.userTower, "Wu", "bu"), (m.itemTower, "Wi", "bi")):
            flat = tw.dense.w.view(-1).cpu()
            o.t[wn].copy_(flat[:E * S].view(E, S)); o.t[bn].copy_(flat[E * S:E * S + S])
    g = torch.Generator(device=dev); g.manual_seed(3)
    uid = torch.randint(2, U + 2, (B,), generator=g, device=dev, dtype=torch.int32)
    iid = torch.randint(2, min(I, 400) + 2, (B,), generator=g, device=dev, dtype=torch.int32)   # repeated items: accidental hits
    lref = o.loss_and_grads(uid.cpu().numpy(), iid.cpu().numpy(), cand_ids=iid.cpu().numpy())
    lgot = m._step(uid, iid, None, True)
    np.testing.assert_allclose(lgot.item(), float(lref), rtol=2e-3)
    worst = (0.0, 0.0, "")
    for got, name in ((m.userTower.emb.g, "Eu"), (m.itemTower.emb.g, "Ei"),
                      (m.userTower.dense.g.view(-1)[:E * S].view(E, S), "Wu"), (m.itemTower.dense.g.view(-1)[:E * S].view(E, S), "Wi"),
                      (m.userTower.dense.g.view(-1)[E * S:E * S + S], "bu")):
        mx, rel, _, _ = _grad_err(got.cpu().numpy(), o.t[name].grad.numpy())
        worst = max(worst, (mx, rel, name))
        assert mx <= 1e-2 and rel <= 1e-2, (name, mx, rel)
    print(f"two-tower TC U={U} B={B}: loss rel err {abs(lgot.item() - float(lref)) / abs(float(lref)):.2e}, worst gradient {worst}")
    # three Adagrad steps at the same shape: the losses keep tracking the oracle
    m.compile("Adagrad", learningRate=0.1)
    m.userTower.emb.g.zero_(); m.itemTower.emb.g.zero_(); m.userTower.dense.g.zero_(); m.itemTower.dense.g.zero_()
    for step in range(3):
        uid = torch.randint(2, U + 2, (B,), generator=g, device=dev, dtype=torch.int32)
        iid = torch.randint(2, I + 2, (B,), generator=g, device=dev, dtype=torch.int32)
        lr_ = o.step(uid.cpu().numpy(), iid.cpu().numpy(), cand_ids=iid.cpu().numpy())
        lg_ = m._step(uid, iid, None, True)
        m.optimizer.apply([m.userTower.emb, m.itemTower.emb], dense=[m.userTower.dense, m.itemTower.dense])
        np.testing.assert_allclose(lg_.item(), lr_, rtol=2e-3)


@pytest.mark.parametrize("E,S,B,gtol", [(128, 128, 1000, 5e-4), (100, 60, 332, 5e-4), (64, 64, 76, 5e-4), (100, 60, 333, 1e-2)])
def test_twotower_one_launch_step_equals_the_multi_kernel_step(dev, monkeypatch, E, S, B, gtol):
    """csrc/twotower_fused.cu (one cooperative launch, scores on chip) against the multi-kernel step it replaces
    (BRK_TT_NO_FUSED=1: gemm_tf32 products + inbatch_softmax_kernel) on the same inputs: same TF32 operand rule, so they
    differ by summation order and by ex2.approx in the softmax only -- loss rtol 1e-5; gradients 5e-4 of the largest entry
    (measured 1e-4: a last-bit difference in P now and then moves an operand across a TF32 truncation boundary).  Batch 333:
    the multi-kernel step takes the fp32 product wherever a leading dimension is not a multiple of 4 (the B x B matrix), the
    one-launch step stays on the tensor cores, so there the two differ by TF32 rounding itself (measured 8e-4..4e-3; bound 1e-2
    as against the oracle).
    Then brk_twotower_train_step: five steps with Adagrad in the same launch against five steps + brk_adagrad_dense."""
    from binrec_b200.twoTower import TwoTowerModel
    U, I = 400, 150                                                     # 150 items: accidental hits in every batch
    g = torch.Generator(device=dev); g.manual_seed(5)
    uid = torch.randint(2, U + 2, (B,), generator=g, device=dev, dtype=torch.int32)
    iid = torch.randint(2, I + 2, (B,), generator=g, device=dev, dtype=torch.int32)
    res = {}
    for path in ("fused", "multi"):
        if path == "multi":
            monkeypatch.setenv("BRK_TT_NO_FUSED", "1")
        else:
            monkeypatch.delenv("BRK_TT_NO_FUSED", raising=False)
        m = TwoTowerModel(E, I, U, "u", "i", list(range(U)), list(range(I)), semb=S, device=dev, tensor_cores=True)
        loss = m._step(uid, iid, None, True).item()
        le = m._step(uid, iid, None, False).item()                      # forward only: same loss, accumulators untouched
        grads = [t.g.clone() for t in (m.userTower.emb, m.itemTower.emb, m.userTower.dense, m.itemTower.dense)]
        touched = [t.touched.clone() for t in (m.userTower.emb, m.itemTower.emb)]
        for t in (m.userTower.emb, m.itemTower.emb, m.userTower.dense, m.itemTower.dense):
            t.g.zero_()
        for t in (m.userTower.emb, m.itemTower.emb):
            t.touched.zero_()
        m.compile("Adagrad", learningRate=0.1)
        losses = []
        for k in range(5):
            uk = torch.roll(uid, k); ik = torch.roll(iid, 2 * k)
            losses.append(m._train_ids(uk, ik, None).item())
        res[path] = (loss, le, grads, touched, losses, [t.w.clone() for t in (m.userTower.emb, m.itemTower.emb, m.userTower.dense,
                                                                              m.itemTower.dense)],
                     [t.g.abs().max().item() for t in (m.userTower.emb, m.itemTower.emb, m.userTower.dense, m.itemTower.dense)],
                     [int(t.touched.count_nonzero().item()) for t in (m.userTower.emb, m.itemTower.emb)])
    f, mk = res["fused"], res["multi"]
    np.testing.assert_allclose(f[0], mk[0], rtol=1e-5)
    np.testing.assert_allclose(f[1], f[0], rtol=1e-6)
    for a, b_ in zip(f[2], mk[2]):
        scale = b_.abs().max().item()
        assert (a - b_).abs().max().item() <= gtol * scale, ((a - b_).abs().max().item(), scale)
    for a, b_ in zip(f[3], mk[3]):
        assert torch.equal(a, b_)                                       # same rows marked
    np.testing.assert_allclose(f[4], mk[4], rtol=2e-5 if gtol < 1e-3 else 1e-4)
    # weights after five steps: Adagrad moves an entry by lr * g / sqrt(0.1 + sum g^2), i.e. an entry whose g is small against
    # the largest one sees the <= 2e-3 * max|g| gradient difference as lr * dg / 0.32 per step: bound 5 steps * 0.1 * 1e-2
    for a, b_ in zip(f[5], mk[5]):
        assert (a - b_).abs().max().item() <= (5e-3 if gtol < 1e-3 else 5e-2)
    assert f[6] == [0.0] * 4 and f[7] == [0, 0] and mk[6] == [0.0] * 4 and mk[7] == [0, 0]   # accumulators and bitmasks left clean
