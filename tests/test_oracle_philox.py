"""Oracle pins: Random123 known-answer vectors for Philox4x32-10 and sampler invariants."""
import numpy as np

from oracle import philox as P

KATS = [  # (counter, key, expected) -- Random123 kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], (0, 0), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, (0xffffffff, 0xffffffff), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], (0xa4093822, 0x299f31d0),
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_philox_known_answers():
    for ctr, key, exp in KATS:
        out = P.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert [int(x) for x in out] == exp


def test_philox_vectorised_equals_scalar():
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2**32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    allv = P.philox4x32_10(ctr, (123, 456))
    for i in range(64):
        assert np.array_equal(allv[i], P.philox4x32_10(ctr[i:i + 1], (123, 456))[0])


def _toy(U=50, I=40, n=600, seed=3):
    rng = np.random.default_rng(seed)
    key = np.unique(rng.integers(0, U * I, n))
    return (key // I).astype(np.int32), (key % I).astype(np.int32)


def test_bpr_negatives_never_positive_and_chunking_invariant():
    U, I = 50, 40
    pu, pi = _toy(U, I)
    indptr, sitems = P.build_csr(pu, pi, U)
    neg = P.bpr_negatives(pu, 7, 3, I, indptr, sitems)
    pos = set(zip(pu.tolist(), pi.tolist()))
    assert all((u, j) not in pos for u, j in zip(pu.tolist(), neg.tolist()))
    assert neg.min() >= 0 and neg.max() < I
    # counter-based: any slice of the stream can be generated independently
    part = P.bpr_negatives(pu[100:200], 7, 3, I, indptr, sitems, first_index=100)
    assert np.array_equal(part, neg[100:200])
    assert not np.array_equal(neg, P.bpr_negatives(pu, 7, 4, I, indptr, sitems))


def test_bpr_negatives_vectorised_form_equals_the_definition():
    U, I = 50, 40
    pu, pi = _toy(U, I)
    # user 3 interacts with everything, user 4 with nothing
    keep = (pu != 3) & (pu != 4)
    pu = np.concatenate([pu[keep], np.full(I, 3, dtype=np.int32)]); pi = np.concatenate([pi[keep], np.arange(I, dtype=np.int32)])
    indptr, sitems = P.build_csr(pu, pi, U)
    q = np.concatenate([pu, np.array([3, 4, 4, 3], dtype=np.int32)])
    assert np.array_equal(P.bpr_negatives(q, 11, 2, I, indptr, sitems, first_index=5),
                          P.bpr_negatives_scalar(q, 11, 2, I, indptr, sitems, first_index=5))


def test_bpr_negatives_are_uniform_over_non_interacted_items():
    # one user with positives {1, 2, 5} of 8 items: every non-interacted item must be reachable, roughly equally often
    I = 8
    pu = np.zeros(3, dtype=np.int32); pi = np.array([1, 2, 5], dtype=np.int32)
    indptr, sitems = P.build_csr(pu, pi, 1)
    neg = P.bpr_negatives(np.zeros(20000, dtype=np.int32), 3, 0, I, indptr, sitems)
    vals, counts = np.unique(neg, return_counts=True)
    assert vals.tolist() == [0, 3, 4, 6, 7]
    assert counts.min() > 3600 and counts.max() < 4400


def test_bpr_negatives_saturated_user_returns_last_candidate():
    U, I = 2, 3
    pu = np.array([0, 0, 0, 1], dtype=np.int32); pi = np.array([0, 1, 2, 0], dtype=np.int32)
    indptr, sitems = P.build_csr(pu, pi, U)
    neg = P.bpr_negatives(np.array([0, 1], dtype=np.int32), 1, 0, I, indptr, sitems)
    assert 0 <= neg[0] < I and neg[1] in (1, 2)


def test_neumf_negatives_follow_marginals():
    pu, pi = _toy()
    nu, ni = P.neumf_negatives(pu, pi, 4 * len(pu), 7, 0)
    assert set(nu.tolist()) <= set(pu.tolist()) and set(ni.tolist()) <= set(pi.tolist())
    nu2, ni2 = P.neumf_negatives(pu, pi, 100, 7, 0, first_index=50)
    assert np.array_equal(nu2, nu[50:150]) and np.array_equal(ni2, ni[50:150])
