"""GPU parity of the ranking-metrics kernel and the topKmetrics mirror against golden vectors
produced by the reference's own topKMetrics / getAverage (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import topk as OT

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "topk_golden.json")))


def test_topkmetrics_matches_reference_golden(dev):
    from binrec_b200 import topKmetrics as TK
    results = []
    for c in GOLD["metrics"]:
        preds = [(u, [(s, i) for s, i in t]) for u, t in c["preds"]]
        pos = [tuple(p) for p in c["pos"]]
        got = TK.topKMetrics(preds, pos, c["users"], c["items"])
        assert got == c["out"]
        results.append(got)
    assert TK.getAverage(results) == GOLD["average_out"]
    preds = [('u1', [(.9, 'i1'), (.8, 'i2')]), ('u2', [(.7, 'i3'), (.6, 'i1')])]
    pos = [('u1', 'i1'), ('u2', 'i2'), ('u2', 'i1')]
    assert TK.topKMetrics(preds, pos, ['u1', 'u2', 'u3'], ['i1', 'i2', 'i3']) == GOLD["survey_example"]


def test_topkmetrics_empty_positives_raise(dev):
    from binrec_b200 import topKmetrics as TK
    with pytest.raises(ZeroDivisionError):
        TK.topKMetrics([("u", [(1.0, "i")])], [], ["u"], ["i"])


def test_metrics_kernel_at_ml1m_shape_with_ndcg(dev):
    from binrec_b200 import topKmetrics as TK
    rng = np.random.default_rng(0)
    U, I, k = 6040, 3706, 10
    ids = np.stack([rng.permutation(I)[:k] for _ in range(U)]).astype(np.int32)
    key = np.unique(rng.integers(0, U * I, 200000))
    pu, pi = key // I, key % I
    tp, hits, ndcg_sum, n_real = TK.topk_counts(torch.from_numpy(ids).to(dev), None, pu, pi, U)
    ref = OT.topk_metrics_arrays(ids, np.arange(U), pu, pi, I)
    assert (tp, n_real - tp, hits / U) == (ref["tp"], ref["fn"], ref["hitRate"])
    assert abs(ndcg_sum / U - ref["ndcg"]) < 1e-12


def test_topkratings_format_and_scores(dev):
    from binrec_b200 import topKmetrics as TK

    class DotModel:
        def __init__(self, uw, iw):
            self.uw, self.iw = uw, iw

        def score_vectors(self, usersId, itemsId):
            return self.uw[torch.as_tensor(usersId, device=dev)], self.iw[torch.as_tensor(itemsId, device=dev)]

    rng = np.random.default_rng(1)
    uw = (rng.integers(-4, 5, size=(20, 16)) / 8.0).astype(np.float32)
    iw = (rng.integers(-4, 5, size=(50, 16)) / 8.0).astype(np.float32)
    users, items = [3, 7, 11], list(range(49, -1, -1))            # catalog order defines the tie rule
    out = TK.topKRatings(5, DotModel(torch.from_numpy(uw).to(dev), torch.from_numpy(iw).to(dev)), users, items)
    for (u, lst), uu in zip(out, users):
        assert u == uu and len(lst) == 5
        scores = [(float(uw[uu] @ iw[i]), i) for i in items]
        ref = OT.topk_insert_reference(scores, 5)
        assert [i for _, i in lst] == [i for _, i in ref]
        assert [float(s) for s, _ in lst] == [s for s, _ in ref]


class _FakeBPR:
    """What full_auc / mean_average_precision_k need from a model: device, get_layer_weights(name)."""
    def __init__(self, P, Q, dev):
        self.device = dev
        self.w = {"user_embedding": torch.from_numpy(np.ascontiguousarray(P, dtype=np.float32)).to(dev),
                  "item_embedding": torch.from_numpy(np.ascontiguousarray(Q, dtype=np.float32)).to(dev)}

    def get_layer_weights(self, name):
        return self.w[name]


def test_full_auc_and_map_match_the_executed_reference(dev):
    """binrec_b200.BPRModel.full_auc / mean_average_precision_k (one SGEMM + one counting kernel per chunk of users)
    against tests/golden/bpr_eval_golden.npz, recorded by EXECUTING the reference's functions (src/models/bpr.py:230-289):
    exact on the exact-arithmetic case with heavy ties, 1e-6 where fp32 summation order may move a near-tie."""
    import json
    import os
    from binrec_b200.BPRModel import full_auc, mean_average_precision_k, _rank_eval
    from oracle import bpr as OB
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bpr_eval_golden.npz"))
    for name in ("random", "ties", "small_k"):
        g = lambda key: G[f"{name}/{key}"]
        items, truth = g("items").tolist(), [(u, t) for u, t in json.loads(str(g("truth")))]
        m = _FakeBPR(g("P"), g("Q"), dev)
        tol = 1e-12 if name == "ties" else 1e-6
        np.testing.assert_allclose(full_auc(m, truth, items), float(g("auc")), rtol=tol)
        for k, want in zip(g("ks").tolist(), g("map")):
            np.testing.assert_allclose(mean_average_precision_k(m, truth, items, k=k), want, rtol=tol)
    # per-user values against the oracle restatement, exact-arithmetic inputs (bit-exact scores, heavy ties)
    rng = np.random.default_rng(9)
    P = (rng.integers(-2, 3, size=(50, 8)) / 2.0).astype(np.float32); Q = (rng.integers(-2, 3, size=(300, 8)) / 2.0).astype(np.float32)
    items = rng.permutation(300)[:257].tolist()
    truth = [(int(u), [int(x) for x in rng.choice(items, size=int(rng.integers(1, 40)), replace=False)]) for u in range(50)]
    auc, ap = _rank_eval(_FakeBPR(P, Q, dev), truth, items, 10)
    pos_of = {it: j for j, it in enumerate(items)}
    for r, (u, t) in enumerate(truth):
        a, p = OB.auc_and_ap_at_k(Q[items] @ P[u], [pos_of[x] for x in t], len(t), 10)
        assert auc[r] == pytest.approx(a, rel=1e-13) and ap[r] == pytest.approx(p, rel=1e-13)


def test_full_auc_and_map_edge_cases(dev):
    from binrec_b200.BPRModel import full_auc, mean_average_precision_k
    m = _FakeBPR(np.eye(3, 4), np.arange(20, dtype=np.float32).reshape(5, 4), dev)
    items = [4, 2, 0, 1]
    with pytest.raises(ValueError):                                         # items.index(p) of an unknown item
        full_auc(m, [(0, [3])], items)
    with pytest.raises(ZeroDivisionError):                                  # nobody has a true item: sum([]) / 0
        full_auc(m, [(0, [])], items)
    with pytest.raises(ZeroDivisionError):                                  # score / min(len(actual), k)
        mean_average_precision_k(m, [(0, [])], items)
    with pytest.raises(ValueError):                                         # every catalog item positive: one class only
        full_auc(m, [(1, [4, 2, 0, 1])], items)
    # user 0 scores items by their first coordinate: item 4 (16) > 2 (8) > 1 (4) > 0 (0); truth {2}: rank 1 of 4
    assert full_auc(m, [(0, [2])], items) == pytest.approx(2.0 / 3.0)
    assert mean_average_precision_k(m, [(0, [2])], items, k=1) == 0.0
    assert mean_average_precision_k(m, [(0, [2])], items, k=2) == pytest.approx(0.5)
    assert mean_average_precision_k(m, [(0, [2, 2, 2])], items, k=100) == pytest.approx(0.5 / 3.0)   # len(actual) counts repeats
