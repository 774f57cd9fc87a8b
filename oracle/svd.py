"""ORACLE (test infrastructure, never the product path): the reference's biased-SVD model
(/root/reference/src/origin_models/svd/SVD.py), SURVEY.md section 8 row f4.

Restated, with the reference's quirks kept on purpose:
  * fit_model (:187-221): ratings are visited one at a time in file order;
        e  = r - (bu + bi + mu + <q_i, p_u>)                                     (:198)
        q' = q + lr * (e * p  - reg * q)                                         (:203)
        p' = p + lr * (e * q' - reg * p)         -- the UPDATED item vector      (:204)
        bu' = bu + lr * (e * bu - breg * bu)     -- error times the bias itself  (:205)
        bi' = bi + lr * (e * bi - breg * bi)                                     (:206)
    all in float64 (NumPy's default for np.random.random / np.zeros, :446-449);
  * predict (:179-185) = bu + bi + mu + <q_i, p_u>; mean_generic_error (:223-247) averages f(error);
  * get_rating / place_in_quintile (:255-270); digest (:105-124): dense ids in order of first appearance and the
    global mean rating (the chunk-wise running mean of :139-161 equals the plain mean up to rounding).

Pinning: **pinned by executed reference code** -- tests/golden/make_svd_golden.py imports SVD.py (TensorFlow / TFRS /
smbclient / git stubbed; they are only touched by do_topk, the SMB reader and get_config) and records inputs and
outputs of digest, fit_model, predict, mean_square_error, mean_absolute_error, recommend and place_in_quintile in
tests/golden/svd_golden.npz; tests/test_oracle_svd.py holds this file to those vectors.

svd_c.c is the same loop in C for sizes Python cannot reach in seconds (the full-size CPU baseline); it is checked
against this module and the golden vectors.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def place_in_quintile(value, quintiles):
    q1, median, q3 = quintiles
    if value > q3:
        return 4
    if value > median:
        return 3
    if value > q1:
        return 2
    return 1


def quintile_rating(transaction_count, quantity_sum, tc_scale=0.5, qs_scale=0.5, tc_quintiles=(1, 2, 4),
                    qs_quintiles=(1, 1, 2)):
    """get_rating with RATING_COLUMN = None (SVD.py:258-262)."""
    tc = np.array([place_in_quintile(v, tc_quintiles) for v in transaction_count], dtype=np.float64)
    qs = np.array([place_in_quintile(v, qs_quintiles) for v in quantity_sum], dtype=np.float64)
    return tc_scale * tc + qs_scale * qs


def digest(raw_users, raw_items, ratings):
    """(user_vocab, item_vocab, users, items, global mean): convert_ids + calculate_average (SVD.py:105-161)."""
    def first_occurrence(a):
        uniq, first, inv = np.unique(a, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty(len(uniq), dtype=np.int64)
        rank[order] = np.arange(len(uniq))
        return uniq[order], rank[inv].astype(np.int32)
    uv, u = first_occurrence(np.asarray(raw_users))
    iv, i = first_occurrence(np.asarray(raw_items))
    return uv, iv, u, i, float(np.mean(np.asarray(ratings, dtype=np.float64)))


def fit_epoch(users, items, ratings, P, Q, bu, bi, mu, lr, emb_reg, bias_reg):
    """One pass of fit_model, in place (float64 arrays), rating by rating."""
    for k in range(len(users)):
        u, i = int(users[k]), int(items[k])
        q, p = Q[i].copy(), P[u].copy()
        e = ratings[k] - (bu[u] + bi[i] + mu + np.dot(q, p))
        q = q + lr * (e * p - emb_reg * q)
        p = p + lr * (e * q - emb_reg * p)
        bu[u] = bu[u] + lr * (e * bu[u] - bias_reg * bu[u])
        bi[i] = bi[i] + lr * (e * bi[i] - bias_reg * bi[i])
        Q[i], P[u] = q, p


def predict(users, items, P, Q, bu, bi, mu):
    return bu[users] + bi[items] + mu + np.einsum("nd,nd->n", Q[items], P[users])


def errors(users, items, ratings, P, Q, bu, bi, mu):
    """(mean squared error, mean absolute error) over a rating set (SVD.py:223-253)."""
    e = np.asarray(ratings, dtype=np.float64) - predict(users, items, P, Q, bu, bi, mu)
    return float(np.mean(e * e)), float(np.mean(np.abs(e)))


def dependency_levels(users, items, num_users, num_items):
    """level[k] = 1 + max(level of the previous rating of the same user, of the same item): ratings of one level touch
    pairwise disjoint rows, and running the levels in order reproduces the sequential pass exactly (what the device
    schedule of csrc/svd.cu relies on)."""
    lu = np.zeros(num_users, dtype=np.int64)
    li = np.zeros(num_items, dtype=np.int64)
    lev = np.empty(len(users), dtype=np.int64)
    for k in range(len(users)):
        l = max(lu[users[k]], li[items[k]]) + 1
        lev[k] = l
        lu[users[k]] = l
        li[items[k]] = l
    return lev


# ---- C port for full sizes ------------------------------------------------------------------------------------
_lib = None


def build_c():
    """gcc -O2 oracle/svd_c.c -> oracle/_build/libsvd_oracle.so (no -ffast-math: same arithmetic as the loop above)."""
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libsvd_oracle.so")
    src = os.path.join(HERE, "svd_c.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", so, src], check=True)
    return so


def c_lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "_build", "libsvd_oracle.so")
        if not os.path.exists(so):
            so = build_c()
        _lib = ctypes.CDLL(so)
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
        _lib.svd_fit_epoch.argtypes = [ip, ip, dp, ctypes.c_int64, dp, dp, dp, dp, ctypes.c_double, ctypes.c_int32,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_double]
        _lib.svd_fit_epoch.restype = None
    return _lib


def fit_epoch_c(users, items, ratings, P, Q, bu, bi, mu, lr, emb_reg, bias_reg):
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    users = np.ascontiguousarray(users, dtype=np.int32); items = np.ascontiguousarray(items, dtype=np.int32)
    ratings = np.ascontiguousarray(ratings, dtype=np.float64)
    for a in (P, Q, bu, bi):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    c_lib().svd_fit_epoch(users.ctypes.data_as(ip), items.ctypes.data_as(ip), ratings.ctypes.data_as(dp), len(users),
                          P.ctypes.data_as(dp), Q.ctypes.data_as(dp), bu.ctypes.data_as(dp), bi.ctypes.data_as(dp),
                          mu, P.shape[1], lr, emb_reg, bias_reg)
