"""Oracle pins for the BPR loss/gradients and the Keras optimizers: hand-computed fp64 cases
(the reference holds no golden vectors for these; SURVEY.md section 8c, pins (2))."""
import math

import numpy as np

from oracle import bpr as B
from oracle import embedding as E


def test_gather_scatter_roundtrip():
    t = np.arange(12, dtype=np.float32).reshape(4, 3)
    ids = np.array([3, 0, 3], dtype=np.int32)
    out = E.gather_rows(t, ids)
    assert np.array_equal(out, t[[3, 0, 3]])
    acc = E.scatter_add_rows(4, ids, out)
    assert np.array_equal(acc[3], 2 * t[3]) and np.array_equal(acc[0], t[0]) and not acc[1].any()


def test_bpr_hand_case_fp64():
    # U=3, I=4, F=2; one triplet (u=1, p=2, n=0)
    user = np.array([[0.1, 0.2], [0.3, -0.4], [0.5, 0.6]], dtype=np.float64)
    item = np.array([[0.2, 0.1], [0.0, 0.3], [-0.5, 0.7], [0.9, 0.9]], dtype=np.float64)
    x = (0.3 * -0.5 + -0.4 * 0.7) - (0.3 * 0.2 + -0.4 * 0.1)   # = -0.43 - 0.02 = -0.45
    s = 1 / (1 + math.exp(-x))
    loss, gu, gi = B.bpr_loss_and_grads(user, item, [1], [2], [0])
    assert abs(loss - (1 - s)) < 1e-15
    g = -s * (1 - s)
    np.testing.assert_allclose(gu[1], g * (item[2] - item[0]), rtol=1e-14)
    np.testing.assert_allclose(gi[2], g * user[1], rtol=1e-14)
    np.testing.assert_allclose(gi[0], -g * user[1], rtol=1e-14)
    assert not gu[0].any() and not gu[2].any() and not gi[1].any() and not gi[3].any()


def test_bpr_grads_match_finite_differences():
    rng = np.random.default_rng(0)
    user = rng.normal(size=(5, 3)); item = rng.normal(size=(6, 3))
    u = np.array([0, 1, 1, 4]); p = np.array([2, 2, 3, 5]); n = np.array([1, 0, 2, 2])
    loss, gu, gi = B.bpr_loss_and_grads(user, item, u, p, n)
    eps = 1e-6
    for (tab, g) in ((user, gu), (item, gi)):
        for r in range(tab.shape[0]):
            for c in range(tab.shape[1]):
                tab[r, c] += eps; lp = B.bpr_loss_and_grads(user, item, u, p, n)[0]
                tab[r, c] -= 2 * eps; lm = B.bpr_loss_and_grads(user, item, u, p, n)[0]
                tab[r, c] += eps
                assert abs((lp - lm) / (2 * eps) - g[r, c]) < 1e-8


def test_keras_adam_first_steps_fp64():
    # Keras: alpha_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= alpha_t*m/(sqrt(v)+eps), eps=1e-7 outside sqrt
    w = np.array([[1.0, -2.0]]); m = np.zeros_like(w); v = np.zeros_like(w)
    g = np.array([[0.5, 0.0]])
    E.adam_dense_keras(w, m, v, g, 1)
    a1 = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(w[0, 0] - (1.0 - a1 * 0.05 / (math.sqrt(0.001 * 0.25) + 1e-7))) < 1e-15
    assert w[0, 1] == -2.0          # zero grad, zero moments: does not move
    # second step with zero gradient: Keras still moves the row on stale momentum (dense-equivalent)
    w1 = w.copy()
    E.adam_dense_keras(w, m, v, np.zeros_like(g), 2)
    assert w[0, 0] < w1[0, 0]
    # lazy Adam leaves an untouched row alone
    w2 = w.copy()
    E.adam_rows_lazy(w, m, v, np.zeros_like(g), np.array([], dtype=np.int64), 3)
    assert np.array_equal(w, w2)


def test_keras_adagrad_fp64():
    w = np.array([[1.0], [2.0]]); acc = np.full_like(w, 0.1)
    g = np.array([[0.3], [0.0]])
    E.adagrad_rows(w, acc, g, [0], lr=0.1)
    assert abs(acc[0, 0] - 0.19) < 1e-15
    assert abs(w[0, 0] - (1.0 - 0.1 * 0.3 / (math.sqrt(0.19) + 1e-7))) < 1e-15
    assert w[1, 0] == 2.0 and acc[1, 0] == 0.1
    # dense pass is identical (zero gradient rows do not move)
    w2 = np.array([[1.0], [2.0]]); acc2 = np.full_like(w2, 0.1)
    E.adagrad_dense(w2, acc2, g, lr=0.1)
    assert np.array_equal(w, w2) and np.array_equal(acc, acc2)


def test_bpr_exhaustive_triplets_match_the_executed_reference(tmp_path):
    """BPRModel.extractPositivesNegatives against tests/golden/bpr_triplets_golden.json, recorded by EXECUTING the
    reference's method (src/models/BPRModel.py:111-119; tests/golden/make_bpr_triplets_golden.py).  Host logic only."""
    import json
    import os
    import numpy as np
    from binrec_b200.BPRModel import BPRModel
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bpr_triplets_golden.json")))
    for case in cases:
        m = BPRModel(workDir=str(tmp_path))
        m._trainDf = (np.asarray(case["users"]), np.asarray(case["items"]))
        m._productIds = list(case["productIds"])
        for customer, want in case["entries"].items():
            got = [[e["CUSTOMER_ID"], int(e["pPRODUCT_ID"]), int(e["nPRODUCT_ID"])] for e in m.extractPositivesNegatives(int(customer))]
            assert got == want, customer


def _bpr_eval_cases():
    import json
    import os
    import numpy as np
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bpr_eval_golden.npz"))
    for name in ("random", "ties", "small_k"):
        g = lambda key: G[f"{name}/{key}"]
        yield name, g("P"), g("Q"), g("items").tolist(), json.loads(str(g("truth"))), float(g("auc")), g("ks").tolist(), g("map"), g("scores_u0")


def test_auc_and_map_oracle_matches_the_executed_reference():
    """oracle/bpr.py:auc_and_ap_at_k against tests/golden/bpr_eval_golden.npz, recorded by EXECUTING the reference's
    full_auc / mean_average_precision_k / bpr_predict (src/models/bpr.py:122-289; make_bpr_eval_golden.py)."""
    import numpy as np
    from oracle import bpr as OB
    for name, P, Q, items, truth, auc, ks, maps, scores_u0 in _bpr_eval_cases():
        pos_of = {it: j for j, it in reversed(list(enumerate(items)))}
        np.testing.assert_allclose(Q[items] @ P[truth[0][0]], scores_u0, rtol=1e-6, atol=1e-7)
        aucs, aps = [], {k: [] for k in ks}
        for u, true_items in truth:
            s = (Q[items].astype(np.float32) @ P[u].astype(np.float32)).astype(np.float32)
            for k in ks:
                a, ap = OB.auc_and_ap_at_k(s, [pos_of[t] for t in true_items], len(true_items), k)
                aps[k].append(ap)
            aucs.append(a)
        # the "ties" case is exact arithmetic, so equal scores are equal in every summation order: exact agreement
        tol = 1e-12 if name == "ties" else 1e-6
        np.testing.assert_allclose(np.mean(aucs), auc, rtol=tol)
        np.testing.assert_allclose([np.mean(aps[k]) for k in ks], maps, rtol=tol)


def test_rating_triplets_follow_the_script_loop():
    """binrec_b200.BPRModel.rating_triplets against the literal loop of src/models/bpr.py:96-107 (the script cannot be
    executed: it downloads MovieLens and trains at import), on seeded ratings with users lacking positives / negatives."""
    import numpy as np
    from binrec_b200.BPRModel import rating_triplets
    rng = np.random.default_rng(4)
    u = rng.integers(0, 12, 300); m = rng.integers(0, 40, 300); r = rng.integers(1, 6, 300)
    r[u == 3] = 5; r[u == 7] = 2                                   # user 3: no negatives, user 7: no positives
    want, without = [], []
    seen = list(dict.fromkeys(u.tolist()))                         # df_train.user_id.unique()
    for user in seen:
        pos = m[(u == user) & (r > 3)]; neg = m[(u == user) & (r <= 3)]
        if len(neg) == 0 or len(pos) == 0:
            without.append(user); continue
        for p in pos:
            for n in neg:
                want.append((user, int(p), int(n)))
    gu, gp, gn, gw = rating_triplets(u, m, r)
    assert list(zip(gu.tolist(), gp.tolist(), gn.tolist())) == want and gw == without and set(without) == {3, 7}
    e = rating_triplets([], [], [])
    assert len(e[0]) == 0 and e[3] == []
