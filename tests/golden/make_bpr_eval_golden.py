"""Generates tests/golden/bpr_eval_golden.npz by EXECUTING the reference's evaluation functions
(/root/reference/src/models/bpr.py: bpr_predict :122-133, full_auc :230-253, mean_average_precision_k :256-289).
bpr.py is a script that trains a Keras model at import, so the three function definitions are lifted out of its syntax
tree and compiled on their own (their code is the reference's, byte for byte; nothing else of the file runs); the
`model` they query is a stand-in exposing get_layer(name).get_weights() over two NumPy matrices.  sklearn is real.
Runs only in the build container; the .npz is committed.      python tests/golden/make_bpr_eval_golden.py
"""
import ast
import json
import os
from collections import OrderedDict
from typing import Dict

import numpy as np
from sklearn.metrics import roc_auc_score

REF = "/root/reference/src/models/bpr.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpr_eval_golden.npz")
WANTED = ("bpr_predict", "full_auc", "mean_average_precision_k")


def load_functions():
    tree = ast.parse(open(REF).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert [n.name for n in body] == list(WANTED), [n.name for n in body]
    ns = {"np": np, "roc_auc_score": roc_auc_score, "OrderedDict": OrderedDict, "Dict": Dict, "Model": object}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


class Layer:
    def __init__(self, w): self.w = w
    def get_weights(self): return [self.w]


class FakeModel:
    def __init__(self, user, item): self.layers = {"user_embedding": Layer(user), "item_embedding": Layer(item)}
    def get_layer(self, name): return self.layers[name]


def main():
    f = load_functions()
    out = {}
    specs = [("random", 40, 60, 16, False), ("ties", 30, 25, 4, True), ("small_k", 25, 130, 8, False)]
    for k_case, (name, U, I, d, ties) in enumerate(specs):
        rng = np.random.default_rng(20261018 + k_case)
        if ties:                                              # exact arithmetic with many equal scores
            P = (rng.integers(-2, 3, size=(U, d)) / 2.0).astype(np.float32)
            Q = (rng.integers(-2, 3, size=(I + 7, d)) / 2.0).astype(np.float32)
        else:
            P = rng.normal(0, 0.3, size=(U, d)).astype(np.float32)
            Q = rng.normal(0, 0.3, size=(I + 7, d)).astype(np.float32)
        items = rng.permutation(I + 7)[:I].tolist()            # the catalog is a subset of the item table, shuffled
        truth = []
        for u in rng.permutation(U)[: U - 3]:
            n_true = int(rng.integers(1, 9))
            truth.append((int(u), [int(x) for x in rng.choice(items, size=n_true, replace=False)]))
        model = FakeModel(P, Q)
        out[f"{name}/P"] = P; out[f"{name}/Q"] = Q
        out[f"{name}/items"] = np.array(items, dtype=np.int64)
        out[f"{name}/truth"] = np.array(json.dumps(truth))
        out[f"{name}/auc"] = np.float64(f["full_auc"](model, truth, items))
        ks = (5, 20, 100)
        out[f"{name}/ks"] = np.array(ks)
        out[f"{name}/map"] = np.array([f["mean_average_precision_k"](model, truth, items, k=k) for k in ks], dtype=np.float64)
        out[f"{name}/scores_u0"] = np.asarray(f["bpr_predict"](model, truth[0][0], items), dtype=np.float32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {n: (float(out[n + "/auc"]), out[n + "/map"].tolist()) for n, *_ in specs})


if __name__ == "__main__":
    main()
