"""Synthetic binary implicit-feedback data of MovieLens-1M shape (BASELINE.json configs[0..2]).

The reference reads proprietary CSVs from an SMB share
(/root/reference/src/models/NeuMFModel.py:21-27, trainers/loadBinaryMovieLens.py:41-62); none of
that data is in its tree, so workloads are generated.  Output schema follows the reference's
frames: int ids, columns CUSTOMER_ID / PRODUCT_ID, every row a positive (rating 1).

Host-side NumPy only: this is input preparation, not the hot path.
"""
import numpy as np

ML1M_USERS = 6040
ML1M_ITEMS = 3706
ML1M_POSITIVES = 1_000_209
DATA_SEED = 20261018


def make_interactions(num_users=ML1M_USERS, num_items=ML1M_ITEMS, num_pos=ML1M_POSITIVES,
                      seed=DATA_SEED, skew=True):
    """Distinct (user, item) positive pairs, int32 arrays of length num_pos, in generation order.

    skew=True draws users ~ floor(U*r^1.5) and items ~ floor(I*r^2) (power-law head like
    MovieLens); skew=False draws both uniformly.
    """
    if num_pos > 0.6 * num_users * num_items:
        raise ValueError("num_pos too dense for rejection sampling")
    rng = np.random.Generator(np.random.Philox(key=seed))
    seen = np.empty(0, dtype=np.int64)
    chunks = []
    have = 0
    while have < num_pos:
        m = int((num_pos - have) * 1.5) + 1024
        ru, ri = rng.random(m), rng.random(m)
        if skew:
            u = np.floor(num_users * ru ** 1.5).astype(np.int64)
            i = np.floor(num_items * ri ** 2.0).astype(np.int64)
        else:
            u = np.floor(num_users * ru).astype(np.int64)
            i = np.floor(num_items * ri).astype(np.int64)
        u = np.minimum(u, num_users - 1)
        i = np.minimum(i, num_items - 1)
        key = u * num_items + i
        # keep first occurrence inside the chunk, drop anything already seen
        _, first = np.unique(key, return_index=True)
        first.sort()
        key = key[first]
        key = key[~np.isin(key, seen, assume_unique=False)]
        chunks.append(key)
        seen = np.concatenate([seen, key])
        have += len(key)
    key = np.concatenate(chunks)[:num_pos]
    return (key // num_items).astype(np.int32), (key % num_items).astype(np.int32)


def build_csr(users, items, num_users):
    """Per-user sorted positive item lists: (indptr int64 [U+1], sorted_items int32 [P]).
    This is the membership structure the device sampler probes (csrc/sampler.cu)."""
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    # distinct (user, item) pairs, sorted by user then item: repeated interactions count once (the sampler's
    # rank-select needs strictly increasing lists)
    width = int(items.max()) + 1 if len(items) else 1
    keys = np.unique(users * width + items)
    su, si = keys // width, keys % width
    counts = np.bincount(su, minlength=num_users)
    indptr = np.zeros(num_users + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, si.astype(np.int32)


def train_test_split(users, items, test_size=0.2, seed=DATA_SEED + 1):
    """Seeded row split standing in for sklearn's unseeded train_test_split
    (/root/reference/src/models/NeuMFModel.py:32)."""
    n = len(users)
    rng = np.random.Generator(np.random.Philox(key=seed))
    perm = rng.permutation(n)
    n_test = int(np.ceil(n * test_size))
    te, tr = perm[:n_test], perm[n_test:]
    return (users[tr], items[tr]), (users[te], items[te])
