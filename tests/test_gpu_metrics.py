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
