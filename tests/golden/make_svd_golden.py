"""Generates tests/golden/svd_golden.npz by EXECUTING the reference's own biased-SVD code
(/root/reference/src/origin_models/svd/SVD.py: digest :105-124, fit_model :187-221, predict :179-185,
mean_square_error / mean_absolute_error :223-253, get_rating / place_in_quintile :255-270, recommend :286-299) on small
seeded frames.  SVD.py imports TensorFlow, TFRS, smbclient and git at module level only for do_topk / the SMB reader /
get_config; they are stubbed here and never reached.  Runs only in the build container (the reference tree is not on
the GPU box); the .npz it writes is committed.      python tests/golden/make_svd_golden.py
"""
import importlib.util
import io
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pandas as pd

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "svd_golden.npz")


def load_svd():
    for name in ("tensorflow", "tensorflow_recommenders", "smbclient", "git", "trainers", "trainers.topKMetrics",
                 "src", "src.AAUfilename", "src.benchmarkLogger"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["src.AAUfilename"].getAAUfilename = lambda p: p
    sys.modules["src.benchmarkLogger"].benchThread = object
    spec = importlib.util.spec_from_file_location("ref_svd", os.path.join(REF, "src/origin_models/svd/SVD.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.VERBOSE = False
    return mod


def frames(rng, U, I, n, n_chunks, rating_kind):
    """Chunks like the reference's five files: raw (non-dense) ids, a skewed item popularity."""
    raw_u = rng.permutation(10 * U)[:U] + 1000
    raw_i = rng.permutation(10 * I)[:I] + 50000
    u = raw_u[np.minimum((U * rng.random(n) ** 1.5).astype(int), U - 1)]
    i = raw_i[np.minimum((I * rng.random(n) ** 2.0).astype(int), I - 1)]
    df = pd.DataFrame({"CUSTOMER_ID": u, "PRODUCT_ID": i})
    if rating_kind == "column":
        df["RATING_TYPE"] = rng.integers(0, 2, n).astype(np.float64)
    elif rating_kind == "stars":
        df["RATING_TYPE"] = rng.integers(1, 6, n).astype(np.float64)
    else:
        df["TRANSACTION_COUNT"] = rng.integers(1, 8, n)
        df["QUANTITY_SUM"] = rng.integers(1, 5, n)
    cuts = np.linspace(0, n, n_chunks + 1).astype(int)
    return [df.iloc[a:b].reset_index(drop=True) for a, b in zip(cuts[:-1], cuts[1:])]


def main():
    svd = load_svd()
    out = {}
    specs = [  # name, U, I, n, chunks, d, epochs, lr, emb_reg, bias_reg, rating kind
        ("tiny", 5, 4, 40, 2, 3, 2, 0.01, 0.0, 0.01, "column"),
        ("defaults", 60, 40, 1500, 5, 50, 1, 0.01, 0.0, 0.01, "column"),
        ("stars_reg", 40, 30, 1200, 3, 8, 3, 0.02, 0.05, 0.02, "stars"),
        ("quintiles", 30, 25, 800, 4, 16, 2, 0.01, 0.01, 0.01, "quintile"),
    ]
    for k, (name, U, I, n, n_chunks, d, epochs, lr, ereg, breg, kind) in enumerate(specs):
        rng = np.random.default_rng(20261018 + k)
        chunks = frames(rng, U, I, n, n_chunks, kind)
        svd.NUMBER_OF_CHUNKS_TO_EAT = n_chunks
        svd.LEARNING_RATE, svd.EMBEDDING_REGULARIZATION, svd.BIAS_REGULARIZATION = lr, ereg, breg
        svd.NUMBER_OF_EMBEDDINGS = d
        svd.RATING_COLUMN = None if kind == "quintile" else "RATING_TYPE"
        with redirect_stdout(io.StringIO()):
            user_ids, item_ids, uid_max, iid_max, mu = svd.digest(chunks)
        nu, ni = uid_max + 1, iid_max + 1
        P = rng.random((nu, d)) * (1 / d)          # SVD.py:446-447
        Q = rng.random((ni, d)) * (1 / d)
        bu = rng.standard_normal(nu) * 0.1         # the reference starts from zero biases (:448-449), where its bias rule
        bi = rng.standard_normal(ni) * 0.1         # (error * bias) keeps them zero for ever; non-zero starts exercise it
        if name == "defaults":
            bu[:] = 0.0; bi[:] = 0.0
        full = pd.concat(chunks, ignore_index=True)
        out[f"{name}/raw_users"] = full["CUSTOMER_ID"].to_numpy(np.int64)
        out[f"{name}/raw_items"] = full["PRODUCT_ID"].to_numpy(np.int64)
        out[f"{name}/ratings"] = np.array([svd.get_rating(row) for _, row in full.iterrows()], dtype=np.float64)
        if kind == "quintile":
            out[f"{name}/transaction_count"] = full["TRANSACTION_COUNT"].to_numpy(np.int64)
            out[f"{name}/quantity_sum"] = full["QUANTITY_SUM"].to_numpy(np.int64)
        out[f"{name}/users"] = np.array([user_ids[x] for x in full["CUSTOMER_ID"]], dtype=np.int32)
        out[f"{name}/items"] = np.array([item_ids[x] for x in full["PRODUCT_ID"]], dtype=np.int32)
        out[f"{name}/user_vocab"] = np.array(list(user_ids.keys()), dtype=np.int64)      # insertion order = dense id order
        out[f"{name}/item_vocab"] = np.array(list(item_ids.keys()), dtype=np.int64)
        out[f"{name}/global_bias"] = np.float64(mu)
        out[f"{name}/hyper"] = np.array([lr, ereg, breg, epochs, d], dtype=np.float64)
        for nm, a in (("P0", P), ("Q0", Q), ("bu0", bu), ("bi0", bi)):
            out[f"{name}/{nm}"] = a.copy()
        mse, mae = [], []
        with redirect_stdout(io.StringIO()):
            for _ in range(epochs):
                svd.fit_model(chunks, P, Q, bu, bi, mu, user_ids, item_ids)
                mse.append(svd.mean_square_error(chunks, P, Q, bu, bi, mu, user_ids, item_ids))
                mae.append(svd.mean_absolute_error(chunks, P, Q, bu, bi, mu, user_ids, item_ids))
        for nm, a in (("P1", P), ("Q1", Q), ("bu1", bu), ("bi1", bi)):
            out[f"{name}/{nm}"] = a.copy()
        out[f"{name}/mse"] = np.array(mse)
        out[f"{name}/mae"] = np.array(mae)
        out[f"{name}/pred0"] = np.float64(svd.predict(int(out[f"{name}/users"][0]), int(out[f"{name}/items"][0]), P, Q, bu, bi, mu))
        rec = svd.recommend(P[0], Q, 3)                 # SVD.py:286-299: three best items of user 0 by dot product
        out[f"{name}/recommend_u0"] = np.array(sorted(((r.prediction, r.index) for r in rec), reverse=True), dtype=np.float64)
    out["quintile_cases"] = np.array([[v, svd.place_in_quintile(v, (1, 2, 4)), svd.place_in_quintile(v, (1, 1, 2))]
                                      for v in range(0, 8)], dtype=np.int64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", {k: float(out[k + "/mse"][-1]) for k in ("tiny", "defaults", "stars_reg", "quintiles")})


def cv_protocol():
    """tests/golden/svd_cv_protocol.json: the chunked-dataset protocol of movielens_cross_validation (SVD.py:301-347)
    and of the main loop (:519-551) recorded from the executed class on a 23-row file: chunk sizes, which chunks the
    iterator yields for every position of the test-set index, the order the folds are visited, the error text."""
    import json
    import tempfile
    svd = load_svd()
    n = 23
    df = pd.DataFrame({"user_id": np.arange(n) + 100, "item_id": np.arange(n) + 500, "rating": (np.arange(n) % 3 == 0) * 1.0})
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "all.csv")
        df.to_csv(path, index=False)
        svd.FILE_PATH = path
        np.random.seed(0)                                   # .sample(frac=1) draws from the global NumPy state
        ds = svd.movielens_cross_validation(path, 5, ["user_id", "item_id", "rating"])
    # np.array_split of a DataFrame gives DataFrames under the pandas the reference was written for and plain arrays
    # under the pandas 3 of this container; the protocol (sizes, indices, iteration) is the same code either way
    users_of = lambda ch: ch["user_id"].to_numpy() if hasattr(ch, "columns") else np.asarray(ch)[:, 0]
    chunk_of = {}
    for c, chunk in enumerate(ds.chunks):
        for uid in users_of(chunk):
            chunk_of[int(uid)] = c
    yielded = lambda: [chunk_of[int(users_of(ch)[0])] for ch in ds]
    out = {"sizes": [len(c) for c in ds.chunks], "initial_test_set_index": ds.test_set_index, "initial_yield": yielded()}
    ds.use_no_test_set()
    out["no_test_set_yield"] = yielded()
    try:
        ds.get_test_set()
    except Exception as e:                                  # noqa: BLE001 -- the reference raises a bare Exception
        out["no_test_set_error"] = str(e)
    folds = []
    while ds.next_cross_validation_distribution():
        test = ds.get_test_set()
        folds.append({"test_set_index": ds.test_set_index, "test_chunk": chunk_of[int(users_of(test)[0])],
                      "test_rows": len(test), "yield": yielded()})
    out["folds"] = folds
    out["final_test_set_index"] = ds.test_set_index
    dst = os.path.join(os.path.dirname(OUT), "svd_cv_protocol.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", dst, out["sizes"], [f["test_set_index"] for f in folds])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cv":
        cv_protocol()
    else:
        main()
