"""Pins oracle/topk.py against golden vectors produced by the reference's own code
(tests/golden/make_golden.py ran /root/reference/trainers/topKmetrics.py)."""
import json
import os

import numpy as np
import pytest

from oracle import topk as T

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "topk_golden.json")))


def test_topk_matches_reference_insertion_with_ties():
    for c in GOLD["topk"]:
        l = [(s, i) for i, s in enumerate(c["scores"])]
        exp = [(s, i) for s, i in c["out"]]
        assert T.topk_insert_reference(l, c["k"]) == exp
        v, ids = T.topk_from_scores(np.array([c["scores"]], dtype=np.float32), c["k"])
        assert ids[0].tolist() == [i for _, i in exp]
        assert v[0].tolist() == [s for s, _ in exp]


def test_metrics_match_reference():
    for c in GOLD["metrics"]:
        preds = [(u, [(s, i) for s, i in t]) for u, t in c["preds"]]
        pos = [tuple(p) for p in c["pos"]]
        assert T.topk_metrics(preds, pos, c["users"], c["items"]) == c["out"]
        # array form: map string ids to ints
        um = {u: j for j, u in enumerate(c["users"])}; im = {i: j for j, i in enumerate(c["items"])}
        ids = np.array([[im[i] for _, i in t] for _, t in preds], dtype=np.int32)
        got = T.topk_metrics_arrays(ids, np.arange(len(c["users"])), [um[u] for u, _ in pos],
                                    [im[i] for _, i in pos], len(c["items"]))
        for key, val in c["out"].items():
            assert got[key] == val, key
        assert 0.0 <= got["ndcg"] <= 1.0


def test_survey_example_and_average():
    assert GOLD["survey_example"] == {"tp": 2, "tn": 4, "fp": 2, "fn": 1, "precision": 0.5,
                                      "recall": 2 / 3, "hitRate": 2 / 3}
    assert T.get_average(GOLD["average"]) == GOLD["average_out"]


def test_empty_positives_raise_like_reference():
    with pytest.raises(ZeroDivisionError):
        T.topk_metrics([("u", [(1.0, "i")])], [], ["u"], ["i"])      # recall = 0/0, as in the reference


def test_ndcg_hand_case():
    # user 0: positives {5, 7}; ranking [7, 1, 5] -> DCG = 1 + 1/log2(4) = 1.5 ; IDCG = 1 + 1/log2(3)
    got = T.topk_metrics_arrays(np.array([[7, 1, 5]]), [0], [0, 0], [5, 7], 10)
    assert abs(got["ndcg"] - 1.5 / (1 + 1 / np.log2(3))) < 1e-12


def test_sharded_merge_equals_unsharded_with_ties():
    rng = np.random.default_rng(0)
    Q = rng.integers(-4, 5, size=(37, 16)).astype(np.float32) / 8
    C = rng.integers(-4, 5, size=(203, 16)).astype(np.float32) / 8
    v, i = T.brute_force_topk(Q, C, 10)
    pv, pi = [], []
    for s in range(0, 203, 64):
        a, b = T.brute_force_topk(Q, C[s:s + 64], 10)
        pv.append(a); pi.append(b + s)
    mv, mi = T.merge_topk(pv, pi, 10)
    assert np.array_equal(mi, i) and np.array_equal(mv, v)


def test_bf16_round_is_rne():
    x = np.array([1.0, 1.00390625, 1.01171875, -2.5, 3.1415927], dtype=np.float32)
    import torch
    assert np.array_equal(T.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


def test_mirror_private_topk_helpers_match_the_executed_reference():
    """binrec_b200.topKmetrics.__topk / __insertSorted (the reference's module-private helpers, kept by name) against
    the golden outputs of the EXECUTED reference __topk (tests/golden/topk_golden.json) and its insertion rule."""
    import json
    import os
    import random
    from binrec_b200 import topKmetrics as T
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "topk_golden.json")))
    topk, insert = getattr(T, "__topk"), getattr(T, "__insertSorted")
    for c in g["topk"]:
        got = topk([(s, i) for i, s in enumerate(c["scores"])], c["k"])
        assert [list(x) for x in got] == c["out"]
    random.seed(0)
    for _ in range(500):                                            # insertion keeps the list descending and stable
        l = sorted([(random.randint(0, 5), j) for j in range(random.randint(1, 8))], key=lambda x: x[0], reverse=True)
        v = (random.randint(0, 6), "new")
        a = list(l); insert(a, v)
        assert [x[0] for x in a] == sorted([x[0] for x in a], reverse=True) and len(a) == len(l) + 1
        assert all(x[0] >= v[0] for x in a[:a.index(v)]) and all(x[0] < v[0] for x in a[a.index(v) + 1:])
