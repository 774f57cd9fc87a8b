"""Generates tests/golden/bpr_triplets_golden.json by EXECUTING the reference's BPRModel.extractPositivesNegatives
(/root/reference/src/models/BPRModel.py:111-119) -- the exhaustive (positive, non-interacted) enumeration behind its
triplet frame.  TensorFlow / Keras / sklearn / RModel are stubbed (module-level imports only; the method is pandas and
plain Python).  Runs only in the build container; the JSON is committed.   python tests/golden/make_bpr_triplets_golden.py
"""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout
from unittest import mock

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bpr_triplets_golden.json")


def load_class():
    for name in ("tensorflow", "tensorflow.keras", "tensorflow.keras.layers", "tensorflow.keras.models",
                 "tensorflow.keras.optimizers", "tensorflow.python", "tensorflow.python.distribute",
                 "tensorflow.python.distribute.distribute_lib", "sklearn", "sklearn.model_selection"):
        sys.modules.setdefault(name, mock.MagicMock())
    sys.modules["tensorflow"].function = lambda f: f
    rm = types.ModuleType("src.models.RModel")
    rm.RModel = type("RModel", (), {"__init__": lambda self, name: None})
    for name, mod in (("src", types.ModuleType("src")), ("src.models", types.ModuleType("src.models")), ("src.models.RModel", rm)):
        sys.modules.setdefault(name, mod)
    spec = importlib.util.spec_from_file_location("ref_bpr", os.path.join(REF, "src/models/BPRModel.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.BPRModel


def main():
    BPR = load_class()
    rng = np.random.default_rng(20261018)
    cases = []
    for U, I, n in ((4, 5, 9), (6, 8, 25), (3, 3, 9)):
        users = rng.integers(0, U, n); items = rng.integers(0, I, n)
        m = BPR.__new__(BPR)
        m._trainDf = pd.DataFrame({"CUSTOMER_ID": users, "PRODUCT_ID": items})
        m._productIds = sorted(set(items.tolist()))
        out = {}
        with redirect_stdout(io.StringIO()):
            for c in range(U + 1):                          # U itself: a customer without rows
                out[str(c)] = [[int(e["CUSTOMER_ID"]), int(e["pPRODUCT_ID"]), int(e["nPRODUCT_ID"])]
                               for e in m.extractPositivesNegatives(c)]
        cases.append({"users": users.tolist(), "items": items.tolist(), "productIds": m._productIds, "entries": out})
    with open(OUT, "w") as f:
        json.dump(cases, f)
    print("wrote", OUT, [sum(len(v) for v in c["entries"].values()) for c in cases])


if __name__ == "__main__":
    main()
