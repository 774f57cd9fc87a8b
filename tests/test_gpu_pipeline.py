"""GPU parity (through the C ABI) of the device input pipeline against
oracle/pipeline.py: everything here is integer / index work, so every comparison is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import philox as PX
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


def PL():
    from binrec_b200 import pipeline
    return pipeline


@pytest.mark.parametrize("n", [1, 2, 5, 306, 4097, 1_000_003])
def test_epoch_permutation_bit_exact(dev, n):
    for seed, epoch, salt in ((7, 0, 0), (2 ** 32 - 1, 2 ** 31 + 3, 1)):
        got = PL().epoch_permutation(n, seed, epoch, salt, device=dev).cpu().numpy()
        assert np.array_equal(got, OP.feistel_perm(n, seed, epoch, salt))
        assert np.array_equal(got, PL().epoch_permutation_host(n, seed, epoch, salt))
    if n > 10:
        got = PL().epoch_permutation(n, 7, 0, 0, first=3, count=5, device=dev).cpu().numpy()
        assert np.array_equal(got, OP.feistel_perm(n, 7, 0, 0)[3:8])


def test_epoch_permutation_errors(dev):
    from binrec_b200._native import BrkError
    assert PL().epoch_permutation(10, 7, 0, 0, first=10, count=0, device=dev).numel() == 0
    with pytest.raises(BrkError):
        PL().epoch_permutation(10, 7, 0, 0, first=8, count=5, device=dev)
    with pytest.raises(BrkError):
        PL().epoch_permutation(0, 7, 0, 0, first=0, count=0, device=dev)


def _toy(P, U, I, seed=1):
    rng = np.random.default_rng(seed)
    key = rng.choice(U * I, P, replace=False)
    return (key // I).astype(np.int32), (key % I).astype(np.int32)


@pytest.mark.parametrize("P,U,I,ratio", [(1, 3, 4, 0), (1, 3, 4, 4), (400, 30, 25, 3), (50_000, 900, 700, 4), (777, 50, 60, 0.5)])
@pytest.mark.parametrize("reject", [False, True])
def test_neumf_epoch_build_bit_exact(dev, P, U, I, ratio, reject):
    pu, pi = _toy(P, U, I)
    n_neg = int(round(P * ratio))
    indptr, sitems = PX.build_csr(pu, pi, U)
    want = OP.neumf_epoch_build(pu, pi, n_neg, 7, 2, reject=reject, indptr=indptr, sorted_items=sitems, num_items=I)
    t = lambda a: torch.from_numpy(a).to(dev)
    got = PL().neumf_epoch_build(t(pu), t(pi), n_neg, 7, 2, reject=reject, csr_indptr=t(indptr), csr_items=t(sitems))
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    if P + n_neg > 20:                                   # a row range equals the slice of the whole frame
        part = PL().neumf_epoch_build(t(pu), t(pi), n_neg, 7, 2, reject=reject, csr_indptr=t(indptr),
                                      csr_items=t(sitems), first=7, count=11)
        for g, w in zip(part, want):
            assert np.array_equal(g.cpu().numpy(), w[7:18])


def test_neumf_epoch_build_argument_errors(dev):
    from binrec_b200._native import BrkError
    pu = torch.zeros(4, dtype=torch.int32, device=dev)
    with pytest.raises(TypeError):
        PL().neumf_epoch_build(pu.long(), pu, 4, 7, 0)
    with pytest.raises(ValueError):
        PL().neumf_epoch_build(pu, pu, 4, 7, 0, reject=True)
    with pytest.raises(BrkError):
        PL().neumf_epoch_build(pu, pu, 4, 7, 0, first=6, count=5)


def test_neumf_epoch_build_full_ml1m_epoch_properties(dev):
    """BASELINE.json configs[0] size: 1 000 209 positives + 4 negatives each = 5 001 045 rows.  Properties that
    need no oracle run: the positives come through as a permutation, the labels count them, with rejection no
    negative is a known positive, and a second epoch differs."""
    from binrec_b200 import synth
    pu, pi = synth.make_interactions()
    P, n_neg = len(pu), 4 * len(pu)
    indptr, sitems = synth.build_csr(pu, pi, synth.ML1M_USERS)
    t = lambda a: torch.from_numpy(a).to(dev)
    u, i, y = PL().neumf_epoch_build(t(pu), t(pi), n_neg, 7, 0, reject=True, csr_indptr=t(indptr), csr_items=t(sitems))
    assert u.numel() == P + n_neg and int(y.sum().item()) == P
    key = u.long() * synth.ML1M_ITEMS + i.long()
    pos_keys = torch.sort(key[y == 1]).values.cpu().numpy()
    assert np.array_equal(pos_keys, np.sort(pu.astype(np.int64) * synth.ML1M_ITEMS + pi))
    neg_keys = key[y == 0].cpu().numpy()
    assert np.isin(neg_keys, pos_keys).sum() <= 2        # survivors of eight draws: (collision rate)^8
    # labels are spread through the frame: every 16 384-row batch holds close to 1/5 positives
    frac = y[:(P + n_neg) // 16384 * 16384].view(-1, 16384).mean(dim=1)
    assert float(frac.min()) > 0.17 and float(frac.max()) < 0.23
    u2, i2, y2 = PL().neumf_epoch_build(t(pu), t(pi), n_neg, 7, 1)
    assert (y2 != y).float().mean().item() > 0.2


def _check_vocab(dev, keys_host, offset):
    v = PL().Vocabulary(dev)
    ids = v.build(keys_host, offset=offset)
    want_ids, want_vocab = OP.factorize_first_occurrence(PL().pack_keys(keys_host), offset)
    assert np.array_equal(ids.cpu().numpy(), want_ids)
    assert v.size == len(want_vocab)
    assert np.array_equal(v.keys.cpu().numpy().view(np.uint64), want_vocab)
    return v


@pytest.mark.parametrize("n,distinct", [(1, 1), (7, 3), (1000, 50), (4096, 4096), (4097, 10), (300_000, 6040), (300_000, 250_000)])
def test_vocabulary_build_matches_first_occurrence_order(dev, n, distinct):
    rng = np.random.default_rng(n + distinct)
    pool = np.unique(rng.integers(-2 ** 40, 2 ** 40, 2 * distinct + 8))
    pool = rng.permutation(pool)[:distinct]
    keys = pool[rng.integers(0, distinct, n)]
    for offset in (0, 2):
        _check_vocab(dev, keys, offset)


def test_vocabulary_strings_lookup_and_oov(dev):
    rng = np.random.default_rng(0)
    col = np.array([str(x) for x in rng.integers(1, 1683, 20_000)])          # ml-100k movie ids as strings
    v = _check_vocab(dev, col, 2)
    assert v.host_keys("S") == list(dict.fromkeys(col.tolist()))               # pd.unique order
    probe = np.array(["1", "1682", "9999", "abc", col[0]])
    got = v.lookup(probe, oov=1).cpu().numpy()
    want = OP.vocab_lookup(PL().pack_keys(probe), PL().pack_keys(np.array(v.host_keys("S"))), offset=2, oov=1)
    assert np.array_equal(got, want)
    assert got[2] == 1 and got[3] == 1 and got[4] == 2
    assert v.lookup(np.array([], dtype=np.int64)).numel() == 0


def test_vocabulary_empty_and_reserved_key(dev):
    v = PL().Vocabulary(dev)
    ids = v.build(np.array([], dtype=np.int64))
    assert ids.numel() == 0 and v.size == 0
    assert v.lookup(np.array([3, 4])).cpu().tolist() == [1, 1]
    with pytest.raises(ValueError):
        v.build(np.array([1, -1]))


def test_vocabulary_20m_keys_properties(dev):
    """BASELINE.json configs[3] scale (20 M user ids): properties only -- ids of first occurrences are
    0,1,2,... in order of appearance, every id maps back to its key, lookups return the same ids."""
    n, distinct = 20_000_000, 3_000_000
    g = torch.Generator(device=dev); g.manual_seed(5)
    raw = torch.randint(0, distinct, (n,), device=dev, generator=g)
    keys = raw * 2654435761 + 12345                                           # spread-out 64-bit patterns
    v = PL().Vocabulary(dev)
    ids = v.build(keys, offset=0).long()
    assert torch.equal(v.keys[ids], keys)
    firsts = torch.zeros(v.size, dtype=torch.int64, device=dev).fill_(n)
    firsts.scatter_reduce_(0, ids, torch.arange(n, device=dev), reduce="amin")
    assert bool((firsts[1:] > firsts[:-1]).all())                              # rank order = first-occurrence order
    assert v.size == int(torch.unique(raw).numel())
    assert torch.equal(v.lookup(keys[:1_000_000]).long(), ids[:1_000_000])


def test_neumf_dataset_uses_the_device_frame(dev):
    """bootstrapDataset -> NeuMFDataset: the frame equals the oracle's, batches are permuted per epoch with the
    keyed batch permutation, iteration yields the reference's ({"user","item"}, label) structure."""
    from binrec_b200.NeuMFModel import NeuMFModel
    pu, pi = _toy(5000, 300, 200)
    m = NeuMFModel()
    ds = m.bootstrapDataset((pu, pi), negRatio=3., batchSize=128, shuffle=True)
    want = OP.neumf_epoch_build(pu, pi, 15000, m.samplerSeed, 0)
    for g, w in zip((ds.u, ds.i, ds.y), want):
        assert np.array_equal(g.cpu().numpy(), w)
    assert len(ds) == (20000 + 127) // 128
    assert np.array_equal(ds.batch_order(3), OP.feistel_perm(len(ds), m.samplerSeed, 3, OP.SALT_BATCHES))
    x, y = next(iter(ds))
    assert set(x) == {"user", "item"} and x["user"].numel() == 128 and y.numel() == 128
    m.rejectCollisions = True
    ds2 = m.bootstrapDataset((pu, pi), negRatio=3., batchSize=128)
    indptr, sitems = PX.build_csr(pu, pi, 300)
    want2 = OP.neumf_epoch_build(pu, pi, 15000, m.samplerSeed, 0, reject=True, indptr=indptr, sorted_items=sitems,
                                 num_items=200)
    for g, w in zip((ds2.u, ds2.i, ds2.y), want2):
        assert np.array_equal(g.cpu().numpy(), w)
