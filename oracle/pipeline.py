"""ORACLE (test infrastructure, never the product path): the device input pipeline of SURVEY.md
section 8 rows f1 / f2 -- epoch permutation, the fused NeuMF epoch builder, and id factorisation.

What it restates
  * /root/reference/src/models/NeuMFModel.py:102-109 (`bootstrapDataset`): positives get label 1,
    `negRatio * P` negatives are rows drawn with replacement from the positive pairs with the item
    column independently re-drawn, label 0, and the merged frame is row-shuffled once
    (`mergeDf.sample(frac=1.)`); `.batch()` then `.shuffle()` permutes whole batches per epoch
    (`:117-121`).  pandas' global RNG is unseeded, so the *order* is defined here ("brk perm v1").
  * `pd.unique` first-occurrence vocabularies (/root/reference/trainers/loadBinaryMovieLens.py:16-19,
    58-61) and the `StringLookup(vocabulary=...)` index rule (/root/reference/trainers/twoTower.py:33-36:
    index 0 = mask, 1 = OOV, vocabulary from 2 in TF 2.3 / 2.4).

Pinning: **parity unpinned at the pandas/TF boundary** (unseeded RNG, TensorFlow not installable);
the factorisation is pinned against `pandas.unique` / `pandas.factorize` executed here
(tests/test_oracle_pipeline.py), the permutation by its defining property (a bijection of [0, n)) and
by the literal per-element definition below.

"brk perm v1": a keyed bijection of [0, n) that needs no sort and no memory -- a six-round Feistel
network over `bits` = max(2, ceil(log2 n)) bits, split into a left part of bits // 2 and a right part of
bits - bits // 2 bits (the widths swap every round, so odd widths work and the domain is < 2n), with
cycle walking: values that land in [n, 2^bits) go through the network again, which keeps the map a
bijection of [0, n).  Round r uses the key pair (k0_r, k1_r) = words 0, 1 of
Philox4x32-10(counter = (r, 0, salt, 0x5E), key = (seed, epoch)) and the round function
F(R) = mix32((R ^ k0_r), k1_r) truncated to the width of the left part, mix32 being the two-multiply
integer finaliser  x ^= x >> 16; x *= 0x7feb352d; x ^= x >> 15; x += k1; x *= 0x846ca68b; x ^= x >> 16.
(Philox supplies the keys, not the per-element rounds: 12 multiplies per evaluation instead of 120, so
the shuffle runs at memory speed instead of being ALU-bound.)
"""
import numpy as np

from . import philox as PX

TAG_PERM = 0x5E
PERM_ROUNDS = 6
SALT_ROWS = 0        # the one-off row shuffle of the merged frame (NeuMFModel.py:109)
SALT_BATCHES = 1     # the per-epoch batch-order shuffle (NeuMFModel.py:120)
MAX_ATTEMPTS = 8     # collision rejection: attempts 0..7, the last one is kept whatever it is


def perm_bits(n):
    bits = 2
    while (1 << bits) < n:
        bits += 1
    return bits


def perm_round_keys(seed, epoch, salt):
    ctr = np.array([[r, 0, salt & 0xFFFFFFFF, TAG_PERM] for r in range(PERM_ROUNDS)], dtype=np.uint32)
    w = PX.philox4x32_10(ctr, (seed, epoch))
    return w[:, 0].astype(np.uint64), w[:, 1].astype(np.uint64)


def _mix32(x, k0, k1):
    m = np.uint64(0xFFFFFFFF)
    x = (x ^ k0) & m
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15)
    x = (x + k1) & m
    x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    return x


def _feistel_once(x, bits, k0, k1):
    hl = bits // 2
    hr = bits - hl
    L = x >> np.uint64(hr)
    R = x & np.uint64((1 << hr) - 1)
    wl, wr = hl, hr
    for r in range(PERM_ROUNDS):
        F = _mix32(R & np.uint64(0xFFFFFFFF), k0[r], k1[r]) & np.uint64((1 << wl) - 1)
        L, R = R, L ^ F
        wl, wr = wr, wl
    return (L << np.uint64(hr)) | R


def feistel_perm(n, seed, epoch, salt=SALT_ROWS, first=0, count=None):
    """perm[j] for j in [first, first+count): int64 array, a bijection of [0, n) over the full range."""
    count = n - first if count is None else count
    bits = perm_bits(n)
    k0, k1 = perm_round_keys(seed, epoch, salt)
    x = np.arange(first, first + count, dtype=np.uint64)
    todo = np.arange(count)
    out = np.empty(count, dtype=np.int64)
    while len(todo):
        x = _feistel_once(x, bits, k0, k1)
        done = x < np.uint64(n)
        out[todo[done]] = x[done].astype(np.int64)
        todo, x = todo[~done], x[~done]
    return out


def feistel_perm_scalar(j, n, seed, epoch, salt=SALT_ROWS):
    """The definition, one element at a time with Python integers (small cases)."""
    bits = perm_bits(n)
    hl = bits // 2
    hr = bits - hl
    k0, k1 = perm_round_keys(seed, epoch, salt)
    x = int(j)
    while True:
        L, R = x >> hr, x & ((1 << hr) - 1)
        wl, wr = hl, hr
        for r in range(PERM_ROUNDS):
            y = (R ^ int(k0[r])) & 0xFFFFFFFF
            y ^= y >> 16
            y = (y * 0x7FEB352D) & 0xFFFFFFFF
            y ^= y >> 15
            y = (y + int(k1[r])) & 0xFFFFFFFF
            y = (y * 0x846CA68B) & 0xFFFFFFFF
            y ^= y >> 16
            L, R = R, L ^ (y & ((1 << wl) - 1))
            wl, wr = wr, wl
        x = (L << hr) | R
        if x < n:
            return x


def _neg_draw(idx, attempt, P, seed, epoch):
    ctr = np.stack([idx & PX.MASK, idx >> np.uint64(32), np.full_like(idx, attempt),
                    np.full_like(idx, PX.TAG_NEUMF)], axis=1).astype(np.uint32)
    w = PX.philox4x32_10(ctr, (seed, epoch))
    return PX._mulshift(w[:, 0], P), PX._mulshift(w[:, 1], P)


def neumf_negatives_rejecting(pos_users, pos_items, n_neg, seed, epoch, indptr, sorted_items, num_items,
                              first_index=0):
    """The "brk sampler v2" NeuMF stream with collision rejection (row f1): attempt a = 0..7 of sample s
    draws Philox(counter = (s, s >> 32, a, 0x4E)); the first attempt whose (user, item) is not a known
    positive wins, attempt 7 is kept regardless.  With no collisions it equals philox.neumf_negatives."""
    pu, pi = np.asarray(pos_users), np.asarray(pos_items)
    P = len(pu)
    idx = np.arange(first_index, first_index + n_neg, dtype=np.uint64)
    u = np.empty(n_neg, dtype=np.int32)
    i = np.empty(n_neg, dtype=np.int32)
    todo = np.arange(n_neg)
    for attempt in range(MAX_ATTEMPTS):
        a, b = _neg_draw(idx[todo], attempt, P, seed, epoch)
        cu, ci = pu[a].astype(np.int32), pi[b].astype(np.int32)
        u[todo], i[todo] = cu, ci
        if attempt == MAX_ATTEMPTS - 1:
            break
        hit = PX._is_positive(indptr, sorted_items, num_items, cu, ci)
        todo = todo[hit]
        if not len(todo):
            break
    return u, i


def neumf_epoch_build(pos_users, pos_items, n_neg, seed, epoch, reject=False, indptr=None, sorted_items=None,
                      num_items=0, first=0, count=None):
    """Rows [first, first+count) of the shuffled training frame: (users int32, items int32, labels f32).
    Row j holds source row s = perm(j) of concat(positives, negatives): the positive pair s if s < P, else
    negative number s - P of the sampler stream, label 0."""
    pu = np.asarray(pos_users, dtype=np.int32)
    pi = np.asarray(pos_items, dtype=np.int32)
    P = len(pu)
    n = P + n_neg
    count = n - first if count is None else count
    src = feistel_perm(n, seed, epoch, SALT_ROWS, first, count)
    if reject:
        nu, ni = neumf_negatives_rejecting(pu, pi, n_neg, seed, epoch, indptr, sorted_items, num_items)
    else:
        nu, ni = PX.neumf_negatives(pu, pi, n_neg, seed, epoch)
    is_pos = src < P
    u = np.where(is_pos, pu[np.minimum(src, P - 1)], nu[np.maximum(src - P, 0)] if n_neg else 0).astype(np.int32)
    i = np.where(is_pos, pi[np.minimum(src, P - 1)], ni[np.maximum(src - P, 0)] if n_neg else 0).astype(np.int32)
    return u, i, is_pos.astype(np.float32)


def factorize_first_occurrence(keys, offset=0):
    """(ids int32 [n], vocab [n_unique]): vocab = distinct keys in order of first appearance (what
    `pd.unique` returns, loadBinaryMovieLens.py:16-19), ids[j] = offset + position of keys[j] in vocab
    (offset 2 = the StringLookup rule of twoTower.py:33-36)."""
    keys = np.asarray(keys)
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # sorted-unique slot -> first-occurrence rank
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[order] = np.arange(len(uniq))
    return (rank[inv] + offset).astype(np.int32), uniq[order]


def vocab_lookup(keys, vocab, offset=2, oov=1):
    """StringLookup of keys against an existing vocabulary: offset + position, `oov` for unknown keys."""
    keys = np.asarray(keys)
    vocab = np.asarray(vocab)
    order = np.argsort(vocab, kind="stable")
    sv = vocab[order]
    pos = np.searchsorted(sv, keys)
    pos = np.minimum(pos, max(len(sv) - 1, 0))
    found = (sv[pos] == keys) if len(sv) else np.zeros(len(keys), dtype=bool)
    return np.where(found, order[pos] + offset if len(sv) else 0, oov).astype(np.int32)


def pack_key_bytes(strings):
    """Exact 64-bit keys for short byte strings (<= 8 bytes): big-endian bytes in the low-order end, the
    length in no separate field -- zero bytes cannot occur in CSV ids, so distinct strings give distinct keys.
    This is how string id columns ('CUSTOMER_ID', 'MATERIAL': loadBinaryMovieLens.py:49) reach the device."""
    out = np.zeros(len(strings), dtype=np.uint64)
    for j, s in enumerate(strings):
        b = s.encode() if isinstance(s, str) else bytes(s)
        if len(b) > 8 or b"\0" in b:
            raise ValueError("id %r does not fit an exact 64-bit key" % (s,))
        out[j] = int.from_bytes(b, "big")
    return out
