"""Generates tests/golden/topk_golden.json by EXECUTING the reference's own functions
(/root/reference/trainers/topKmetrics.py and src/origin_models/svd/topKMetrics.py) with a stub
`tensorflow` module.  Runs only in the build container (the reference tree is not on the GPU box);
the JSON it writes is committed.   python tests/golden/make_golden.py
"""
import importlib.util
import json
import os
import random
import sys
import types

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "topk_golden.json")


def load(path, name):
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    tk = load(os.path.join(REF, "trainers/topKmetrics.py"), "ref_topk")
    svd = load(os.path.join(REF, "src/origin_models/svd/topKMetrics.py"), "ref_svd_topk")
    ref_topk = getattr(tk, "__topk")
    rnd = random.Random(20261018)
    cases = {"topk": [], "metrics": [], "average": []}
    # __topk: scores on a coarse grid => many ties; k >= 2 (k == 1 raises IndexError in the reference)
    for _ in range(60):
        n = rnd.randint(2, 40)
        k = rnd.randint(2, min(n, 12))
        levels = rnd.choice([3, 5, 9, 1000])
        l = [(rnd.randint(-levels, levels) / 8.0, i) for i in range(n)]
        out = ref_topk(list(l), k)
        cases["topk"].append({"scores": [s for s, _ in l], "k": k, "out": [[s, i] for s, i in out]})
    # topKMetrics on random predictions / positives
    for _ in range(25):
        U, I, k = rnd.randint(2, 12), rnd.randint(4, 15), rnd.randint(1, 4)
        users = ["u%d" % i for i in range(U)]; items = ["i%d" % i for i in range(I)]
        preds = [(u, [(rnd.random(), it) for it in rnd.sample(items, k)]) for u in users]
        pos = list({(rnd.choice(users), rnd.choice(items)) for _ in range(rnd.randint(1, 3 * U))})
        res = tk.topKMetrics(preds, pos, users, items)
        cases["metrics"].append({"users": users, "items": items,
                                 "preds": [[u, [[s, i] for s, i in t]] for u, t in preds],
                                 "pos": [list(p) for p in pos], "out": res})
        cases["average"].append(res)
    cases["average_out"] = svd.getAverage(cases["average"])
    # the worked example quoted in SURVEY.md section 8c
    preds = [('u1', [(.9, 'i1'), (.8, 'i2')]), ('u2', [(.7, 'i3'), (.6, 'i1')])]
    pos = [('u1', 'i1'), ('u2', 'i2'), ('u2', 'i1')]
    cases["survey_example"] = tk.topKMetrics(preds, pos, ['u1', 'u2', 'u3'], ['i1', 'i2', 'i3'])
    with open(OUT, "w") as f:
        json.dump(cases, f)
    print("wrote", OUT, {k: (len(v) if isinstance(v, list) else "1") for k, v in cases.items()})


if __name__ == "__main__":
    main()
