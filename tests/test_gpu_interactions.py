"""GPU side of the interaction formats: CSV -> device factorisation -> cache -> model,
and the device StringLookup of the two-tower mirror.  Index work: bit-exact against pandas / the host dictionary."""
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_csv_to_cache_factorises_like_pd_unique(dev, tmp_path):
    from binrec_b200 import interactions as IX
    rng = np.random.default_rng(4)
    users = np.array(["%04d" % x for x in rng.integers(0, 500, 20_000)])
    mats = np.array(["M%05d" % x for x in rng.integers(0, 3000, 20_000)])
    csv = tmp_path / "tt.csv"
    pd.DataFrame({"CUSTOMER_ID": users, "MATERIAL": mats}).to_csv(csv, index=False)
    c = IX.csv_to_cache(str(csv), str(tmp_path / "tt.brkc"), "twotower", device=dev)
    codes_u, uniq_u = pd.factorize(users)
    codes_m, uniq_m = pd.factorize(mats)
    assert np.array_equal(c.columns["user"], codes_u) and np.array_equal(c.columns["item"], codes_m)
    assert c.vocabulary("user") == uniq_u.tolist() and c.vocabulary("item") == uniq_m.tolist()     # pd.unique order
    assert c.attrs["num_user"] == len(uniq_u) and c.attrs["num_item"] == len(uniq_m)
    d = c.to_device(["user", "item"], device=dev, rowLimit=1000)
    assert d["user"].dtype == torch.int32 and np.array_equal(d["user"].cpu().numpy(), codes_u[:1000])


def test_ml100k_shape_file_and_bpr_training_from_cache(dev, tmp_path):
    """u.data layout (loadBinaryMovieLens.py:8-21) -> cache with dense ids -> BPRModel.train on the cache path."""
    from binrec_b200 import interactions as IX
    from binrec_b200.BPRModel import BPRModel
    rng = np.random.default_rng(7)
    key = rng.choice(943 * 1682, 30_000, replace=False)
    u, m = key // 1682 + 1, key % 1682 + 1
    f = tmp_path / "u.data"
    pd.DataFrame({"u": u, "m": m, "r": rng.integers(1, 6, len(u)), "t": 88}).to_csv(f, sep="\t", header=False, index=False)
    c = IX.csv_to_cache(str(f), str(tmp_path / "ml.brkc"), "ml-100k", device=dev)
    assert len(c) == 30_000 and c.vocabulary("user")[:3] == [str(x) for x in u[:3]]
    assert np.array_equal(c.columns["value"], np.loadtxt(f, usecols=2, dtype=np.float32))
    model = BPRModel(workDir=str(tmp_path))
    model.epochs, model.batchSize, model.numFactor = 2, 1024, 16
    out = model.train(str(tmp_path / "ml.brkc"), None)
    assert out["result"] == "completed"


def test_string_lookup_on_device_equals_host_dictionary(dev):
    from binrec_b200.twoTower import StringLookup
    rng = np.random.default_rng(1)
    vocab = list(dict.fromkeys("%05d" % x for x in rng.integers(0, 5000, 4000)))
    sl = StringLookup(vocab)
    probe = ["%05d" % x for x in rng.integers(0, 6000, 10_000)] + ["", "zz"]
    got = sl.lookup_device(probe, dev).cpu().numpy()
    assert np.array_equal(got, sl(probe))
    assert (got == 1).any() and got.min() >= 1
    # ids without an exact 64-bit key keep the host dictionary
    long_vocab = ["customer-%012d" % j for j in range(50)]
    sl2 = StringLookup(long_vocab)
    assert np.array_equal(sl2.lookup_device(long_vocab[::-1] + ["nobody"], dev).cpu().numpy(), sl2(long_vocab[::-1] + ["nobody"]))
    # integer ids
    sl3 = StringLookup([40, 10, 30])
    assert sl3.lookup_device([10, 99, 40, 30], dev).cpu().tolist() == [3, 1, 2, 4]
    with pytest.raises(ValueError):
        StringLookup(["a", "b", "a"]).lookup_device(["a"], dev)
