"""CPU checks of the input-pipeline oracle (oracle/pipeline.py) and of the library's host-side permutation:
the factorisation is pinned against pandas (`pd.unique` / `pd.factorize`, what the reference calls at
trainers/loadBinaryMovieLens.py:16-19,58-61) executed here; the permutation by its definition and properties."""
import numpy as np
import pandas as pd
import pytest

from oracle import philox as PX
from oracle import pipeline as OP


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 16, 17, 306, 1000, 4097])
def test_feistel_perm_is_a_bijection_and_matches_the_scalar_definition(n):
    for seed, epoch, salt in ((7, 0, 0), (7, 1, 0), (7, 0, 1), (2 ** 32 - 1, 2 ** 31 + 3, 5)):
        p = OP.feistel_perm(n, seed, epoch, salt)
        assert sorted(p.tolist()) == list(range(n))
        if n <= 306:
            assert [OP.feistel_perm_scalar(j, n, seed, epoch, salt) for j in range(n)] == p.tolist()
    # a range of the permutation equals the slice of the whole
    if n > 4:
        assert np.array_equal(OP.feistel_perm(n, 7, 0, 0, first=2, count=n - 3), OP.feistel_perm(n, 7, 0, 0)[2:n - 1])


def test_feistel_perm_depends_on_every_key_word_and_mixes():
    n = 100_000
    base = OP.feistel_perm(n, 7, 0, 0)
    for other in (OP.feistel_perm(n, 8, 0, 0), OP.feistel_perm(n, 7, 1, 0), OP.feistel_perm(n, 7, 0, 1)):
        assert (other != base).mean() > 0.99
    assert (base == np.arange(n)).sum() <= 8                      # expected 1 fixed point
    assert abs(np.corrcoef(base, np.arange(n))[0, 1]) < 0.02
    # successive source rows land far apart (no visible stride)
    assert np.median(np.abs(np.diff(base))) > n / 8


@pytest.mark.parametrize("n", [1, 7, 1000, 65537])
def test_library_host_permutation_equals_the_oracle(n):
    from binrec_b200 import pipeline as PL
    for seed, epoch, salt in ((7, 0, 1), (123456789, 9, 0)):
        assert np.array_equal(PL.epoch_permutation_host(n, seed, epoch, salt), OP.feistel_perm(n, seed, epoch, salt))
    if n > 5:
        assert np.array_equal(PL.epoch_permutation_host(n, 7, 0, 1, first=3, count=2),
                              OP.feistel_perm(n, 7, 0, 1)[3:5])
    from binrec_b200._native import BrkError
    with pytest.raises(BrkError):
        PL.epoch_permutation_host(n, 7, 0, 1, first=n, count=1)


def test_factorize_matches_pandas_unique_and_factorize():
    rng = np.random.default_rng(3)
    for keys in (rng.integers(0, 50, 1000), rng.integers(-2 ** 40, 2 ** 40, 500), np.array([5]), np.array([9, 9, 9]),
                 np.array([str(x) for x in rng.integers(0, 300, 2000)])):
        ids, vocab = OP.factorize_first_occurrence(keys)
        codes, uniques = pd.factorize(keys)
        assert np.array_equal(ids, codes)
        assert np.array_equal(vocab, pd.unique(keys))
        assert np.array_equal(vocab, uniques)
        ids2, _ = OP.factorize_first_occurrence(keys, offset=2)
        assert np.array_equal(ids2, codes + 2)


def test_vocab_lookup_string_lookup_rule():
    vocab = np.array([40, 10, 30])
    got = OP.vocab_lookup(np.array([10, 99, 40, 30, -1]), vocab, offset=2, oov=1)
    assert got.tolist() == [3, 1, 2, 4, 1]
    assert OP.vocab_lookup(np.array([1, 2]), np.array([], dtype=np.int64)).tolist() == [1, 1]


def test_pack_keys_exact_and_agrees_with_the_oracle():
    from binrec_b200 import pipeline as PL
    col = ["1", "22", "1682", "A1B2C3D4", "", "z"]
    k = PL.pack_keys(col)
    assert np.array_equal(k, OP.pack_key_bytes(col))
    assert len(set(k.tolist())) == len(col)
    assert PL.unpack_keys(k, "S") == col
    ints = np.array([0, 1, -2, 2 ** 62], dtype=np.int64)
    assert np.array_equal(PL.unpack_keys(PL.pack_keys(ints), "i"), ints)
    with pytest.raises(ValueError):
        PL.pack_keys(["123456789"])
    with pytest.raises(ValueError):
        PL.pack_keys(np.array([-1]))


def _toy(P=400, U=30, I=25, seed=1):
    rng = np.random.default_rng(seed)
    key = rng.choice(U * I, P, replace=False)
    return (key // I).astype(np.int32), (key % I).astype(np.int32), U, I


def test_epoch_build_is_a_shuffle_of_positives_and_the_sampler_stream():
    pu, pi, U, I = _toy()
    P, n_neg = len(pu), 1200
    u, i, y = OP.neumf_epoch_build(pu, pi, n_neg, 7, 3)
    assert len(u) == P + n_neg and y.sum() == P
    pos = sorted(zip(u[y == 1].tolist(), i[y == 1].tolist()))
    assert pos == sorted(zip(pu.tolist(), pi.tolist()))
    nu, ni = PX.neumf_negatives(pu, pi, n_neg, 7, 3)
    assert sorted(zip(u[y == 0].tolist(), i[y == 0].tolist())) == sorted(zip(nu.tolist(), ni.tolist()))
    # ranges compose
    a = OP.neumf_epoch_build(pu, pi, n_neg, 7, 3, first=100, count=50)
    assert all(np.array_equal(x, z[100:150]) for x, z in zip(a, (u, i, y)))


def test_rejecting_sampler_avoids_known_positives_and_keeps_clean_draws():
    pu, pi, U, I = _toy(P=400, U=30, I=25)            # density 0.53: collisions are common
    indptr, sitems = PX.build_csr(pu, pi, U)
    n_neg = 3000
    nu0, ni0 = PX.neumf_negatives(pu, pi, n_neg, 7, 0)
    nu, ni = OP.neumf_negatives_rejecting(pu, pi, n_neg, 7, 0, indptr, sitems, I)
    hit0 = PX._is_positive(indptr, sitems, I, nu0, ni0)
    hit = PX._is_positive(indptr, sitems, I, nu, ni)
    assert hit0.mean() > 0.3
    assert hit.mean() < 0.02                             # (collision rate)^8 survives
    assert np.array_equal(nu[~hit0], nu0[~hit0]) and np.array_equal(ni[~hit0], ni0[~hit0])
    u, i, y = OP.neumf_epoch_build(pu, pi, n_neg, 7, 0, reject=True, indptr=indptr, sorted_items=sitems, num_items=I)
    assert sorted(zip(u[y == 0].tolist(), i[y == 0].tolist())) == sorted(zip(nu.tolist(), ni.tolist()))
