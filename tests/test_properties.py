"""Property tests (hypothesis) of the oracle and the host-side routing logic -- SURVEY.md section 4 (iv):
scatter-add == dense one-hot product, sharded top-K merge == unsharded top-K, row-shard arithmetic is a bijection,
the SVD tickets are the ordinal within user / item, the epoch permutation is a bijection.  CPU only."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import embedding as OE
from oracle import pipeline as OPL
from oracle import svd as OS
from oracle import topk as OT

SETTINGS = dict(max_examples=60, deadline=None)


@settings(**SETTINGS)
@given(st.integers(1, 40), st.integers(1, 9), st.integers(0, 200), st.integers(0, 2 ** 31 - 1))
def test_scatter_add_is_the_one_hot_product(rows, d, n, seed):
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, rows, n)
    vals = (rng.integers(-8, 9, size=(n, d)) / 8.0).astype(np.float32)          # exact sums
    got = OE.scatter_add_rows(rows, ids, vals)
    onehot = np.zeros((n, rows), np.float32); onehot[np.arange(n), ids] = 1.0
    assert np.array_equal(got, onehot.T @ vals)
    # gather is the adjoint: <gather(T, ids), V> == <T, scatter(ids, V)>
    T = (rng.integers(-8, 9, size=(rows, d)) / 8.0).astype(np.float32)
    assert float((OE.gather_rows(T, ids) * vals).sum()) == float((T * got).sum())


@settings(**SETTINGS)
@given(st.integers(1, 12), st.integers(1, 60), st.integers(1, 4), st.integers(1, 10), st.integers(0, 2 ** 31 - 1))
def test_sharded_topk_merge_equals_unsharded(U, I, shards, k, seed):
    rng = np.random.default_rng(seed)
    k = min(k, I)
    S = (rng.integers(-4, 5, size=(U, I)) / 4.0).astype(np.float32)             # many ties
    want_v, want_i = OT.topk_from_scores(S, k)
    cuts = np.linspace(0, I, shards + 1).astype(int)
    pv, pi = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        kk = min(k, b - a)
        v = np.full((U, k), -np.inf, np.float32); ix = np.full((U, k), -1, np.int64)
        if kk:
            sv, si = OT.topk_from_scores(S[:, a:b], kk)
            v[:, :kk] = sv; ix[:, :kk] = si + a
        pv.append(v); pi.append(ix)
    got_v, got_i = OT.merge_topk(np.stack(pv), np.stack(pi), k)
    assert np.array_equal(got_v, want_v) and np.array_equal(got_i, want_i)


@settings(**SETTINGS)
@given(st.integers(1, 200), st.integers(1, 8), st.integers(0, 2 ** 31 - 1))
def test_row_shard_arithmetic_is_a_bijection(num_rows, G, seed):
    from binrec_b200 import distributed as D
    ids = torch.arange(num_rows)
    own, loc = D.owner_of(ids, G), D.local_row(ids, G)
    assert torch.equal(loc * G + own, ids)                                       # (owner, local row) -> id
    for r in range(G):
        assert int((own == r).sum()) == D.shard_rows(num_rows, r, G)
        assert sorted(loc[own == r].tolist()) == list(range(D.shard_rows(num_rows, r, G)))
    perm, counts = D.bucket_by_owner(torch.from_numpy(np.random.default_rng(seed).integers(0, num_rows, 50)), G)
    assert int(counts.sum()) == 50 and sorted(perm.tolist()) == list(range(50))
    los = [D.local_slice(num_rows, r, G) for r in range(G)]
    assert los[0][0] == 0 and los[-1][1] == num_rows and all(a[1] == b[0] for a, b in zip(los[:-1], los[1:]))


@settings(**SETTINGS)
@given(st.integers(1, 30), st.integers(1, 30), st.integers(0, 300), st.integers(0, 2 ** 31 - 1))
def test_svd_dependency_levels_order_every_conflict(U, I, n, seed):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, n); i = rng.integers(0, I, n)
    lev = OS.dependency_levels(u, i, U, I)
    last_u, last_i = {}, {}
    for k in range(n):                                                           # a rating sits above both predecessors
        want = 1 + max(last_u.get(u[k], 0), last_i.get(i[k], 0))
        assert lev[k] == want
        last_u[u[k]] = last_i[i[k]] = lev[k]
    for l in np.unique(lev):                                                     # one level touches disjoint rows
        idx = np.nonzero(lev == l)[0]
        assert len(set(u[idx])) == len(idx) and len(set(i[idx])) == len(idx)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 3000), st.integers(0, 2 ** 32 - 1), st.integers(0, 50), st.integers(0, 3))
def test_epoch_permutation_is_a_bijection(n, seed, epoch, salt):
    p = OPL.feistel_perm(n, seed, epoch, salt)
    assert sorted(np.asarray(p).tolist()) == list(range(n))


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 400), st.integers(1, 50), st.integers(0, 2 ** 31 - 1))
def test_first_occurrence_factorisation_is_pd_unique_order(n, distinct, seed):
    import pandas as pd
    keys = np.random.default_rng(seed).integers(-distinct, distinct, n)
    ids, vocab = OPL.factorize_first_occurrence(keys)[:2]
    assert np.array_equal(np.asarray(vocab), pd.unique(keys)) and np.array_equal(np.asarray(vocab)[np.asarray(ids)], keys)
