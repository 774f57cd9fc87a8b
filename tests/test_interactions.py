"""Host-side checks of the interaction formats: the binary columnar cache round trip,
its failure modes, and the CSV schemas of the reference (NeuMFModel.py:21-27, NFC_plain.py:72,
loadBinaryMovieLens.py:8-21,41-62).  No GPU: id factorisation on the device is covered in test_gpu_pipeline.py
and test_gpu_interactions.py."""
import io
import os

import numpy as np
import pytest

from binrec_b200 import interactions as IX


def test_cache_round_trip_and_alignment(tmp_path):
    rng = np.random.default_rng(0)
    cols = {"user": rng.integers(0, 6040, 10_001).astype(np.int32), "item": rng.integers(0, 3706, 10_001).astype(np.int32),
            "value": rng.random(10_001).astype(np.float32)}
    vocab = {"user": (np.arange(6040, dtype=np.uint64) * 7, "i")}
    p = str(tmp_path / "a.brkc")
    IX.write_cache(p, cols, vocab, attrs={"schema": "neumf", "num_user": 6040})
    c = IX.InteractionCache(p)
    assert len(c) == 10_001 and c.attrs["num_user"] == 6040
    for k in cols:
        assert c.columns[k].dtype == cols[k].dtype and np.array_equal(c.columns[k], cols[k])
        assert c.columns[k].offset % IX.ALIGN == 0
    assert c.roles == {"user": "id", "item": "id", "value": "value"}
    assert np.array_equal(c.vocabulary("user"), np.arange(6040) * 7)
    assert not os.path.exists(p + f".tmp{os.getpid()}")


def test_cache_empty_and_string_vocabulary(tmp_path):
    from binrec_b200 import pipeline as PL
    p = str(tmp_path / "e.brkc")
    IX.write_cache(p, {"user": np.zeros(0, np.int32), "item": np.zeros(0, np.int32)},
                   {"item": (PL.pack_keys(["A1", "77", "x"]), "S")})
    c = IX.InteractionCache(p)
    assert len(c) == 0 and c.columns["user"].shape == (0,)
    assert c.vocabulary("item") == ["A1", "77", "x"]


def test_cache_rejects_bad_input_and_damaged_files(tmp_path):
    p = str(tmp_path / "b.brkc")
    with pytest.raises(IX.InteractionError):
        IX.write_cache(p, {})
    with pytest.raises(IX.InteractionError):
        IX.write_cache(p, {"user": np.zeros(3, np.int32), "item": np.zeros(4, np.int32)})
    with pytest.raises(IX.InteractionError):
        IX.write_cache(p, {"user": np.array(["a", "b"])})
    IX.write_cache(p, {"user": np.arange(5000, dtype=np.int32)})
    raw = open(p, "rb").read()
    bad = str(tmp_path / "bad.brkc")
    open(bad, "wb").write(b"NOTBRKC\0" + raw[8:])
    with pytest.raises(IX.InteractionError, match="magic"):
        IX.InteractionCache(bad)
    open(bad, "wb").write(raw[:-100])
    with pytest.raises(IX.InteractionError, match="truncated"):
        IX.InteractionCache(bad)
    open(bad, "wb").write(raw[:10])
    with pytest.raises(IX.InteractionError, match="header"):
        IX.InteractionCache(bad)
    hdr = bytearray(raw); hdr[8] = 9
    open(bad, "wb").write(bytes(hdr))
    with pytest.raises(IX.InteractionError, match="version"):
        IX.InteractionCache(bad)


def test_reference_csv_schemas_parse():
    neumf = "CUSTOMER_ID,PRODUCT_ID,MATERIAL,QUANTITY\n3,7,100200,1\n0,2,100300,4\n3,2,100300,1\n"
    c = IX.read_csv_columns(io.StringIO(neumf), "neumf")
    assert c["user"].tolist() == [3, 0, 3] and c["item"].tolist() == [7, 2, 2] and "value" not in c
    assert IX.read_csv_columns(io.StringIO(neumf), "neumf", rowLimit=2)["user"].tolist() == [3, 0]
    ncf = "a,b,c,d,e\n900017,4,55001,9,1\n900018,5,55002,3,0\n"
    c = IX.read_csv_columns(io.StringIO(ncf), "ncf")
    assert c["user"].tolist() == [4, 5] and c["item"].tolist() == [9, 3] and c["value"].tolist() == [1.0, 0.0]
    tt = "CUSTOMER_ID,MATERIAL\n0012,00A7\n0013,00A7\n"
    c = IX.read_csv_columns(io.StringIO(tt), "twotower")
    assert c["user"].tolist() == ["0012", "0013"] and c["item"].tolist() == ["00A7", "00A7"]     # strings keep zeros
    rz = "h1,h2,h3,h4,h5\n0012,1,00A7,5,1\n0013,2,00A8,6,0\n"
    c = IX.read_csv_columns(io.StringIO(rz), "twotower-rdzero")
    assert c["item"].tolist() == ["00A7", "00A8"] and c["value"].tolist() == [1.0, 0.0]
    ml = "196\t242\t3\t881250949\n186\t302\t3\t891717742\n"
    c = IX.read_csv_columns(io.StringIO(ml), "ml-100k")
    assert c["user"].tolist() == ["196", "186"] and c["item"].tolist() == ["242", "302"]
    with pytest.raises(IX.InteractionError):
        IX.read_csv_columns(io.StringIO(neumf), "nope")
    with pytest.raises(IX.InteractionError):
        IX.read_csv_columns(io.StringIO("A,B\n1,2\n"), "neumf")


def test_csv_to_cache_with_raw_integer_ids(tmp_path):
    csv = tmp_path / "t.csv"
    csv.write_text("CUSTOMER_ID,PRODUCT_ID,MATERIAL,QUANTITY\n3,7,1,1\n0,2,1,4\n3,2,1,1\n")
    c = IX.csv_to_cache(str(csv), str(tmp_path / "t.brkc"), "neumf")        # raw ids: no device needed
    assert c.columns["user"].tolist() == [3, 0, 3] and c.attrs["num_user"] == 4 and c.attrs["num_item"] == 8
    csv.write_text("CUSTOMER_ID,PRODUCT_ID,MATERIAL,QUANTITY\n-3,7,1,1\n")
    with pytest.raises(IX.InteractionError):
        IX.csv_to_cache(str(csv), str(tmp_path / "t2.brkc"), "neumf")


def test_loaders_match_the_executed_reference():
    """tests/golden/loader_golden.json holds what the reference's own loaders (trainers/loadBinaryMovieLens.py gfData
    :41-62, movieLensData :8-39) returned for three small files (tests/golden/make_loader_golden.py): parsed id columns
    (strings, leading zeros kept, first row dropped) and the pd.unique vocabularies in first-appearance order."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_golden.json")))

    def first_appearance(col):
        seen = {}
        for x in col:
            seen.setdefault(x, len(seen))
        return list(seen)

    for key, schema, vocab_u, vocab_i in (("gf", "twotower", "usersId", "materialsId"),
                                          ("gf_rdzero", "twotower-rdzero", "usersId", "materialsId"),
                                          ("ml100k", "ml-100k", "usersId", "moviesId")):
        c = IX.read_csv_columns(io.StringIO(g["inputs"][key]), schema)
        want = g[key]
        assert [str(x) for x in c["user"]] == want["users"] and [str(x) for x in c["item"]] == want["items"], key
        assert first_appearance([str(x) for x in c["user"]]) == want[vocab_u], key
        assert first_appearance([str(x) for x in c["item"]]) == want[vocab_i], key
        assert len(want[vocab_u]) == want["nbrUser"]
        if "values" in want:
            assert c["value"].tolist() == want["values"]
    # ml-100k: the reference overwrites every rating with ratedVal (:15); the raw stars stay available here
    assert IX.read_csv_columns(io.StringIO(g["inputs"]["ml100k"]), "ml-100k")["value"].tolist() == [3.0, 3.0, 1.0, 5.0, 2.0]


def test_loader_mirror_returns_the_reference_dicts(tmp_path):
    """binrec_b200.loadBinaryMovieLens.gfData / movieLensData against the executed reference loaders' results."""
    import json
    import os
    from binrec_b200 import loadBinaryMovieLens as LB
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_golden.json")))
    for key, rd in (("gf", False), ("gf_rdzero", True)):
        f = tmp_path / (key + ".csv"); f.write_text(g["inputs"][key])
        r = LB.gfData(str(f), "user", "password", rdZero=rd)
        want = g[key]
        assert r["ratings"]["CUSTOMER_ID"] == want["users"] and r["ratings"]["MATERIAL"] == want["items"]
        assert r["usersId"].tolist() == want["usersId"] and r["materialsId"].tolist() == want["materialsId"]
        assert (r["nbrUser"], r["nbrMaterial"]) == (want["nbrUser"], want["nbrMaterial"])
        if rd:
            assert r["ratings"]["RATING_TYPE"] == want["values"]
    f = tmp_path / "u.data"; f.write_text(g["inputs"]["ml100k"])
    r = LB.movieLensData(1, 0, 0.0, path=str(f))
    want = g["ml100k"]
    assert r["ratings"]["user_id"] == want["users"] and r["ratings"]["movie_id"] == want["items"]
    assert r["ratings"]["rating"] == want["ratings"] and r["usersId"].tolist() == want["usersId"]
    assert r["moviesId"].tolist() == want["moviesId"] and (r["nbrUser"], r["nbrMovie"]) == (want["nbrUser"], want["nbrMovie"])
    assert r["realRat"] == set(zip(want["users"], want["items"]))
