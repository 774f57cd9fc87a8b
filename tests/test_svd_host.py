"""Host-side logic of the biased-SVD mirror that needs no GPU: the chunked-dataset protocol of the reference
(movielens_cross_validation / grundfos_network_drive_files, src/origin_models/svd/SVD.py:301-409) and the fold
distribution of cross_validate over ranks (gloo, world size 2)."""
import numpy as np
import pytest

from test_distributed_cpu import _spawn


def _ds(n=23, chunks=5, **kw):
    from binrec_b200 import SVD as S
    return S, S.RatingChunks.split(np.arange(n) + 100, np.arange(n) + 500, (np.arange(n) % 3 == 0).astype(float), chunks, **kw)


def test_chunk_protocol_follows_the_reference():
    S, ds = _ds()
    assert [len(c[0]) for c in ds.chunks] == [5, 5, 5, 4, 4]                 # np.array_split, SVD.py:310
    assert ds.test_set_index == 4 and len(list(ds)) == 4                      # last chunk held out at first (:306)
    ds.use_no_test_set()
    assert len(list(ds)) == 5                                                 # iterator returns every chunk (:313-314)
    with pytest.raises(Exception, match="There is no test set"):
        ds.get_test_set()                                                     # (:328-330)
    seen = []
    while ds.next_cross_validation_distribution():                           # n-1, ..., 0 then False (:318-324)
        seen.append(ds.test_set_index)
        held = ds.get_test_set()
        assert np.array_equal(held[0], ds.chunks[ds.test_set_index][0])
        assert [c[0][0] for c in ds] == [ds.chunks[c][0][0] for c in range(5) if c != ds.test_set_index]
    assert seen == [4, 3, 2, 1, 0] and ds.test_set_index == -1


def test_chunk_protocol_matches_the_executed_reference_class():
    """tests/golden/svd_cv_protocol.json was recorded by EXECUTING movielens_cross_validation (SVD.py:301-347) and the
    main loop's fold walk (:540-551) on a 23-row file (tests/golden/make_svd_golden.py cv)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "svd_cv_protocol.json")))
    S, ds = _ds()
    chunk_of = {int(u): c for c, ch in enumerate(ds.chunks) for u in ch[0]}
    yielded = lambda: [chunk_of[int(ch[0][0])] for ch in ds]
    assert [len(c[0]) for c in ds.chunks] == g["sizes"]
    assert ds.test_set_index == g["initial_test_set_index"] and yielded() == g["initial_yield"]
    ds.use_no_test_set()
    assert yielded() == g["no_test_set_yield"]
    with pytest.raises(Exception) as e:
        ds.get_test_set()
    assert str(e.value) == g["no_test_set_error"]
    folds = []
    while ds.next_cross_validation_distribution():
        test = ds.get_test_set()
        folds.append({"test_set_index": ds.test_set_index, "test_chunk": chunk_of[int(test[0][0])], "test_rows": len(test[0]),
                      "yield": yielded()})
    assert folds == g["folds"] and ds.test_set_index == g["final_test_set_index"]


def test_grundfos_test_set_keeps_rating_one_rows_only():
    S, ds = _ds(test_positive_only=True)
    u, i, r = ds.get_test_set()                                               # query("RATING_TYPE==1"), SVD.py:389-392
    assert (r == 1).all() and len(u) == int((ds.chunks[4][2] == 1).sum())


def test_frames_need_digest_and_shuffle_is_seeded():
    S, ds = _ds()
    with pytest.raises(S.N.BrkError, match="digest"):
        ds.train_frame()
    a = S.RatingChunks.split(np.arange(50), np.arange(50), np.ones(50), 5, shuffle_seed=3)
    b = S.RatingChunks.split(np.arange(50), np.arange(50), np.ones(50), 5, shuffle_seed=3)
    assert all(np.array_equal(x[0], y[0]) for x, y in zip(a.chunks, b.chunks))
    assert not np.array_equal(np.concatenate([c[0] for c in a.chunks]), np.arange(50))


def test_fold_assignment_covers_every_fold_once():
    from binrec_b200 import SVD as S
    for world in (1, 2, 3, 8):
        got = sorted(f for r in range(world) for f in S.fold_assignment(5, r, world))
        assert got == [0, 1, 2, 3, 4]


def _merge(rank, world):
    from binrec_b200 import SVD as S
    local = {f: {"mse": 10.0 * f + rank} for f in S.fold_assignment(5, rank, world)}
    merged = S.merge_fold_results(local, 5)
    return {f: merged[f]["mse"] for f in sorted(merged)}


def test_fold_results_are_gathered_over_ranks_gloo():
    out = _spawn(_merge, world=2)
    want = {0: 0.0, 1: 11.0, 2: 20.0, 3: 31.0, 4: 40.0}                        # fold f ran on rank f % 2
    assert out[0] == want and out[1] == want


def test_convert_ids_running_average_and_idset_follow_the_reference():
    """convert_ids (:126-137), calculate_average (:139-161), get_idset (:410-416) against the golden digest of the
    executed reference (tests/golden/svd_golden.npz): chunk by chunk they reproduce its vocabularies and global mean."""
    import os
    from binrec_b200 import SVD as S
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "svd_golden.npz"))
    for name, n_chunks in (("tiny", 2), ("stars_reg", 3)):
        g = lambda k: G[f"{name}/{k}"]
        u, i, r = g("raw_users"), g("raw_items"), g("ratings")
        cuts = np.linspace(0, len(u), n_chunks + 1).astype(int)
        user_ids, item_ids, nu, ni, total, avg = {}, {}, 0, 0, 0, 0
        chunks = [(u[a:b], i[a:b], r[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
        for c, chunk in enumerate(chunks):
            nu, ni = S.convert_ids(chunk, c, user_ids, item_ids, nu, ni)
            total, avg = S.calculate_average(chunk, c, total, avg)
        assert list(user_ids) == g("user_vocab").tolist() and list(item_ids) == g("item_vocab").tolist()
        assert (nu, ni) == (len(user_ids), len(item_ids)) and total == len(u)
        np.testing.assert_allclose(avg, g("global_bias"), rtol=1e-13)
        assert S.get_idset(chunks) == set(zip(u.tolist(), i.tolist()))


def test_file_backed_dataset_classes_keep_the_reference_names(tmp_path):
    """movielens_cross_validation / grundfos_network_drive_files / get_data / get_config (SVD.py:79-103, 301-409, 499-514)
    over local CSV files: chunking, held-out chunk, rating == 1 filter of the grundfos test set."""
    import pandas as pd
    from binrec_b200 import SVD as S
    n = 23
    pd.DataFrame({"user_id": np.arange(n) + 100, "item_id": np.arange(n) + 500, "rating": (np.arange(n) % 3 == 0) * 1.0,
                  "other": 0}).to_csv(tmp_path / "all.csv", index=False)
    ml = S.movielens_cross_validation(str(tmp_path / "all.csv"), 5, ["user_id", "item_id", "rating"], shuffle_seed=None)
    assert [len(c[0]) for c in ml.chunks] == [5, 5, 5, 4, 4] and ml.test_set_index == 4
    assert np.array_equal(np.concatenate([c[0] for c in ml.chunks]), np.arange(n) + 100)      # file order without a seed
    assert [len(c[0]) for c in ml] == [5, 5, 5, 4] and [len(c[0]) for c in ml] == [5, 5, 5, 4]  # a new `for` restarts it
    for k in range(1, 4):
        pd.DataFrame({"CUSTOMER_ID": np.arange(6) + 10 * k, "PRODUCT_ID": np.arange(6), "RATING_TYPE": [1, 0, 1, 0, 1, 1]}
                     ).to_csv(tmp_path / f"part_{k}.csv", index=False)
    old = S.NUMBER_OF_FILES
    try:
        S.NUMBER_OF_FILES = 3
        gf = S.get_data(str(tmp_path / "part_{0}.csv"))
    finally:
        S.NUMBER_OF_FILES = old
    assert isinstance(gf, S.grundfos_network_drive_files) and gf.number_of_chunks == 3 and gf.test_set_index == 2
    u, i, r = gf.get_test_set()
    assert (r == 1).all() and len(u) == 4                                                   # query("RATING_TYPE==1")
    with pytest.raises(ValueError):
        S.get_data()
    cfg = S.get_config()
    assert cfg["epochs"] == S.EPOCHS and cfg["learning_rate"] == S.LEARNING_RATE and "git_commit_sha" in cfg
