"""ORACLE (test infrastructure, never the product path): Philox4x32-10 and the negative samplers.

The reference has no counter-based sampler: NeuMF draws negatives with pandas' global RNG
(/root/reference/src/models/NeuMFModel.py:104-105,109) and BPR enumerates every
(positive, non-interacted) pair (/root/reference/src/models/BPRModel.py:111-119,
/root/reference/src/models/bpr.py:96-107).  BASELINE.json's north_star asks for Philox
sampling that is reproducible bit-for-bit, so the stream is *defined here* and the CUDA
kernel (csrc/sampler.cu) must reproduce it exactly.

Pinning: Philox4x32-10 is the Random123 generator (Salmon et al., SC'11); its published
known-answer vectors are checked in tests/test_oracle_philox.py.

Stream definition ("brk sampler v2"):
  key      = (seed & 0xffffffff, epoch & 0xffffffff)
  counter  = (idx & 0xffffffff, idx >> 32, attempt, stream_tag)
  a draw r in [0, 2^32) maps to an index in [0, n) by the multiply-shift  (r * n) >> 32.

  BPR negatives (stream_tag 0xB9): one Philox call per sample (attempt = 0, word 0).  For user
  u with sorted positive list a[0..n) the draw picks rank r = (w0 * (I - n)) >> 32 among the
  I - n NON-interacted items and the negative is the r-th such item in ascending id order:
  j = r + t with t the smallest index in [0, n] such that t == n or a[t] - t > r.  Uniform over
  the non-interacted items (what the reference's exhaustive enumeration samples from,
  BPRModel.py:116), one binary search per sample and no rejection loop (v1 re-drew on a
  collision: same distribution, but the slowest of a batch's samples needed ~10 searches and
  a whole step waited for it).  A user who interacted with every item (n == I) gets
  j = (w0 * I) >> 32, the only case in which the "negative" is a positive.

  NeuMF negatives (stream_tag 0x4E): one Philox call per negative, attempt = 0;
  user = pos_user[(w0 * P) >> 32], item = pos_item[(w1 * P) >> 32].  Like the reference
  (NeuMFModel.py:104-105) user and item follow the empirical popularity marginals and no
  collision check against true positives is made.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

TAG_BPR = 0xB9
TAG_NEUMF = 0x4E


def philox4x32_10(ctr, key):
    """ctr: uint32 [N,4]; key: (k0,k1) ints (scalars or [N] arrays). Returns uint32 [N,4]."""
    ctr = np.asarray(ctr, dtype=np.uint32).reshape(-1, 4)
    c0 = ctr[:, 0].astype(np.uint64)
    c1 = ctr[:, 1].astype(np.uint64)
    c2 = ctr[:, 2].astype(np.uint64)
    c3 = ctr[:, 3].astype(np.uint64)
    k0 = np.broadcast_to(np.asarray(key[0], dtype=np.uint64) & MASK, c0.shape).copy()
    k1 = np.broadcast_to(np.asarray(key[1], dtype=np.uint64) & MASK, c0.shape).copy()
    for _ in range(10):
        p0 = M0 * c0          # 32x32 -> 64, no overflow in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.uint32)


def _mulshift(r, n):
    return ((r.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def build_csr(users, items, num_users):
    """Sorted per-user lists of DISTINCT positives: (indptr int64 [U+1], sorted_items int32 [<= P])."""
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    pairs = np.unique(np.stack([users, items], axis=1), axis=0) if len(users) else np.zeros((0, 2), dtype=np.int64)
    su, si = pairs[:, 0], pairs[:, 1]
    indptr = np.zeros(num_users + 1, dtype=np.int64)
    np.add.at(indptr, su + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, si.astype(np.int32)


def _is_positive(indptr, sorted_items, num_items, u, j):
    """Vectorised membership test via the global sorted key u*I + j."""
    keys = np.repeat(np.arange(len(indptr) - 1, dtype=np.int64), np.diff(indptr)) * num_items \
        + sorted_items.astype(np.int64)
    q = u.astype(np.int64) * num_items + j.astype(np.int64)
    pos = np.searchsorted(keys, q)
    pos = np.minimum(pos, len(keys) - 1) if len(keys) else pos
    return (keys[pos] == q) if len(keys) else np.zeros(len(q), dtype=bool)


def _bpr_draws(users, seed, epoch, first_index):
    n = len(users)
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    ctr = np.stack([(idx & MASK), (idx >> np.uint64(32)), np.zeros(n, dtype=np.uint64),
                    np.full(n, TAG_BPR, dtype=np.uint64)], axis=1).astype(np.uint32)
    return philox4x32_10(ctr, (seed, epoch))[:, 0]


def bpr_negatives_scalar(users, seed, epoch, num_items, indptr, sorted_items, first_index=0):
    """The definition, sample by sample (small inputs; bpr_negatives is the vectorised form)."""
    users = np.asarray(users, dtype=np.int64)
    w0 = _bpr_draws(users, seed, epoch, first_index)
    indptr = np.asarray(indptr, dtype=np.int64)
    sorted_items = np.asarray(sorted_items, dtype=np.int64)
    out = np.empty(len(users), dtype=np.int64)
    for s in range(len(users)):
        lo, hi = int(indptr[users[s]]), int(indptr[users[s] + 1])
        missing = [j for j in range(num_items) if j not in set(sorted_items[lo:hi].tolist())]
        if not missing:
            out[s] = (int(w0[s]) * num_items) >> 32
        else:
            out[s] = missing[(int(w0[s]) * len(missing)) >> 32]
    return out.astype(np.int32)


def bpr_negatives(users, seed, epoch, num_items, indptr, sorted_items, first_index=0):
    """Negatives for samples first_index .. first_index+len(users)-1 (int32 [N])."""
    users = np.asarray(users, dtype=np.int64)
    n = len(users)
    w0 = _bpr_draws(users, seed, epoch, first_index).astype(np.uint64)
    indptr = np.asarray(indptr, dtype=np.int64)
    items = np.asarray(sorted_items, dtype=np.int64)
    base = indptr[users]
    cnt = indptr[users + 1] - base
    full = cnt >= num_items
    r = ((w0 * (np.uint64(num_items) - np.minimum(cnt, num_items).astype(np.uint64))) >> np.uint64(32)).astype(np.int64)
    # smallest t in [0, cnt] with t == cnt or a[t] - t > r  (a[t] - t is non-decreasing), all samples at once
    lo = np.zeros(n, dtype=np.int64)
    hi = cnt.copy()
    while True:
        act = lo < hi
        if not act.any():
            break
        mid = (lo + hi) >> 1
        probe = np.where(act, base + np.minimum(mid, np.maximum(cnt - 1, 0)), 0)
        above = (items[probe] - mid > r) if len(items) else np.zeros(n, dtype=bool)
        hi = np.where(act & above, mid, hi)
        lo = np.where(act & ~above, mid + 1, lo)
    out = r + lo
    out[full] = ((w0[full] * np.uint64(num_items)) >> np.uint64(32)).astype(np.int64)
    return out.astype(np.int32)


def neumf_negatives(pos_users, pos_items, n_neg, seed, epoch, first_index=0):
    """(neg_users int32 [n_neg], neg_items int32 [n_neg]) drawn from the positive pairs."""
    P = len(pos_users)
    idx = np.arange(first_index, first_index + n_neg, dtype=np.uint64)
    ctr = np.stack([(idx & MASK), (idx >> np.uint64(32)), np.zeros(n_neg, dtype=np.uint64),
                    np.full(n_neg, TAG_NEUMF, dtype=np.uint64)], axis=1).astype(np.uint32)
    words = philox4x32_10(ctr, (seed, epoch))
    a = _mulshift(words[:, 0], P)
    b = _mulshift(words[:, 1], P)
    return (np.asarray(pos_users)[a].astype(np.int32), np.asarray(pos_items)[b].astype(np.int32))


def epoch_permutation_keys(n, seed, epoch):
    """uint32 sort keys that define the epoch's row order (stream_tag 0x5F): argsort(stable) of
    word0 of philox(ctr=(i,0,0,0x5F)). The reference shuffles with TF's unseeded RNG
    (NeuMFModel.py:109,120; RModel.py:130), so the order is defined here instead."""
    idx = np.arange(n, dtype=np.uint64)
    ctr = np.stack([(idx & MASK), (idx >> np.uint64(32)), np.zeros(n, dtype=np.uint64),
                    np.full(n, 0x5F, dtype=np.uint64)], axis=1).astype(np.uint32)
    return philox4x32_10(ctr, (seed, epoch))[:, 0]
