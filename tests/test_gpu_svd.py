"""Biased-SVD on the device (csrc/svd.cu through binrec_b200.SVD) against
  * the golden vectors produced by EXECUTING the reference's SVD.py (tests/golden/svd_golden.npz), and
  * the oracle (oracle/svd.py, oracle/svd_c.c) on seeded inputs up to the full ML-1M shape.
float64 everywhere; the only freedom against the reference is the summation order inside the dot product (a lane
butterfly here, BLAS ddot there), hence rtol 1e-10 on parameters / errors -- and bit-equality wherever no dot
product is involved or the order is the same (schedule, repeated runs, different residency)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "svd_golden.npz"))
CASES = ["tiny", "defaults", "stars_reg", "quintiles"]
RTOL = 1e-10


@pytest.fixture(scope="module")
def S():
    from binrec_b200 import SVD
    return SVD


def _case(name):
    g = lambda k: G[f"{name}/{k}"]
    lr, ereg, breg, epochs, d = g("hyper")
    return g, float(lr), float(ereg), float(breg), int(epochs), int(d)


def _dev(*arrays):
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


def _tickets(keys):
    """ordinal of every entry among the entries with the same key, in file order (the definition)."""
    seen = {}
    out = np.empty(len(keys), dtype=np.int32)
    for k, v in enumerate(keys.tolist()):
        out[k] = seen.get(v, 0)
        seen[v] = out[k] + 1
    return out


@pytest.mark.parametrize("name", CASES)
def test_fit_predict_errors_match_the_reference_golden(S, name):
    g, lr, ereg, breg, epochs, d = _case(name)
    P, Q, bu, bi = _dev(g("P0"), g("Q0"), g("bu0"), g("bi0"))
    mu = float(g("global_bias"))
    frame = S.Ratings(g("users"), g("items"), g("ratings"), num_users=P.shape[0], num_items=Q.shape[0])
    for e in range(epochs):
        S.fit_model(frame, P, Q, bu, bi, mu, learning_rate=lr, embedding_regularization=ereg, bias_regularization=breg)
        np.testing.assert_allclose(S.mean_square_error(frame, P, Q, bu, bi, mu), g("mse")[e], rtol=RTOL)
        np.testing.assert_allclose(S.mean_absolute_error(frame, P, Q, bu, bi, mu), g("mae")[e], rtol=RTOL)
    S.check_fit(frame)
    # mean_generic_error (SVD.py:223-247): abs / square are reduced on the device, any other callable on the host
    np.testing.assert_allclose(S.mean_generic_error(abs, frame, P, Q, bu, bi, mu), g("mae")[-1], rtol=RTOL)
    np.testing.assert_allclose(S.mean_generic_error(lambda x: x ** 2, frame, P, Q, bu, bi, mu), g("mse")[-1], rtol=RTOL)
    cubes = (g("ratings") - S.predict(g("users"), g("items"), P, Q, bu, bi, mu).cpu().numpy())
    np.testing.assert_allclose(S.mean_generic_error(lambda x: abs(x) ** 3, frame, P, Q, bu, bi, mu), np.mean(np.abs(cubes) ** 3), rtol=1e-12)
    for got, key in ((P, "P1"), (Q, "Q1"), (bu, "bu1"), (bi, "bi1")):
        np.testing.assert_allclose(got.cpu().numpy(), g(key), rtol=RTOL, atol=1e-14, err_msg=key)
    pred0 = S.predict(int(g("users")[0]), int(g("items")[0]), P, Q, bu, bi, mu)
    np.testing.assert_allclose(pred0, g("pred0"), rtol=RTOL)
    rec = S.recommend(P[0], Q, 3)
    assert [r.index for r in rec] == g("recommend_u0")[:, 1].astype(int).tolist()            # rating_prediction objects
    np.testing.assert_allclose([v for _, v in rec], g("recommend_u0")[:, 0], rtol=RTOL)
    assert str(rec[0]) == f"({rec[0].prediction} @ {rec[0].index})"
    # svd_prediction_doer (SVD.py:163-177): the same prediction, queried with raw ids
    uv, iv = g("user_vocab"), g("item_vocab")
    doer = S.svd_prediction_doer({int(x): j for j, x in enumerate(uv)}, {int(x): j for j, x in enumerate(iv)}, P, Q, bu, bi, mu)
    np.testing.assert_allclose(doer.predict([[int(g("raw_users")[0])], [int(g("raw_items")[0])]]), g("pred0"), rtol=RTOL)


@pytest.mark.parametrize("name", CASES)
def test_digest_matches_the_reference_golden(S, name):
    g, *_ = _case(name)
    user_ids, item_ids, uid_max, iid_max, mu, frame = S.digest(g("raw_users"), g("raw_items"), g("ratings"))
    assert list(user_ids.keys()) == g("user_vocab").tolist() and list(user_ids.values()) == list(range(len(user_ids)))
    assert list(item_ids.keys()) == g("item_vocab").tolist()
    assert uid_max == len(g("user_vocab")) - 1 and iid_max == len(g("item_vocab")) - 1
    assert np.array_equal(frame.users.cpu().numpy(), g("users")) and np.array_equal(frame.items.cpu().numpy(), g("items"))
    np.testing.assert_allclose(mu, g("global_bias"), rtol=1e-13)


def test_quintile_ratings_match_the_reference_golden(S):
    g, *_ = _case("quintiles")
    got = S.get_rating(g("transaction_count").astype(np.float64), g("quantity_sum").astype(np.float64))
    assert np.array_equal(got.cpu().numpy(), g("ratings"))
    for v, a, b in G["quintile_cases"]:
        assert S.place_in_quintile(v, (1, 2, 4)) == a and S.place_in_quintile(v, (1, 1, 2)) == b
    edge = np.array([0.0, 1.0, 1.5, 2.0, 2.5, 4.0, 4.5, 100.0])
    from oracle import svd as OS
    assert np.array_equal(S.get_rating(edge, edge[::-1].copy()).cpu().numpy(), OS.quintile_rating(edge, edge[::-1]))


@pytest.mark.parametrize("n,U,I", [(1, 1, 1), (5000, 1, 1), (20000, 300, 200), (20000, 7, 5000), (100000, 50000, 9)])
def test_schedule_is_the_ordinal_within_user_and_item(S, n, U, I):
    rng = np.random.default_rng(n + U)
    u = np.minimum((U * rng.random(n) ** 1.5).astype(np.int32), U - 1)
    i = np.minimum((I * rng.random(n) ** 2.0).astype(np.int32), I - 1)
    frame = S.Ratings(u, i, np.zeros(n), num_users=U, num_items=I)
    sched = frame.sched.cpu().numpy()[:n]
    assert np.array_equal(sched[:, 0], u) and np.array_equal(sched[:, 1], i)
    assert np.array_equal(sched[:, 2], _tickets(u)) and np.array_equal(sched[:, 3], _tickets(i))


def test_bad_ids_and_shapes_raise(S):
    with pytest.raises(IndexError):
        S.Ratings(np.array([0, 3], np.int32), np.array([0, 0], np.int32), np.zeros(2), num_users=3, num_items=1)
    with pytest.raises(IndexError):
        S.Ratings(np.array([0, -1], np.int32), np.array([0, 0], np.int32), np.zeros(2), num_users=3, num_items=1)
    with pytest.raises(ValueError):
        S.Ratings(np.zeros(2, np.int32), np.zeros(3, np.int32), np.zeros(2))
    frame = S.Ratings(np.array([0, 1], np.int32), np.array([0, 1], np.int32), np.ones(2))
    P, Q, bu, bi = S.init_parameters(2, 2, 4, seed=0)
    with pytest.raises(IndexError):
        S.fit_model(frame, P[:1].contiguous(), Q, bu, bi, 0.5)
    with pytest.raises(TypeError):
        S.fit_model(frame, P.float(), Q, bu, bi, 0.5)
    with pytest.raises(IndexError):
        S.predict(np.array([2], np.int32), np.array([0], np.int32), P, Q, bu, bi, 0.5)
    empty = S.Ratings(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), num_users=2, num_items=2)
    before = P.clone()
    S.fit_model(empty, P, Q, bu, bi, 0.5)                                   # an empty file is a no-op
    assert torch.equal(P, before)
    with pytest.raises(ZeroDivisionError):
        S.mean_square_error(empty, P, Q, bu, bi, 0.5)


@pytest.mark.parametrize("d", [1, 2, 31, 32, 33, 50, 64, 65, 128, 200, 512])
def test_every_row_width_against_the_oracle(S, d):
    from oracle import svd as OS
    rng = np.random.default_rng(d)
    U, I, n = 40, 25, 3000
    u = rng.integers(0, U, n).astype(np.int32); i = np.minimum((I * rng.random(n) ** 2).astype(np.int32), I - 1)
    r = rng.integers(1, 6, n).astype(np.float64)
    P = rng.random((U, d)) / d; Q = rng.random((I, d)) / d
    bu = rng.normal(0, 0.1, U); bi = rng.normal(0, 0.1, I)
    Pd, Qd, bud, bid = _dev(P, Q, bu, bi)
    frame = S.Ratings(u, i, r, num_users=U, num_items=I)
    for _ in range(2):
        OS.fit_epoch_c(u, i, r, P, Q, bu, bi, 3.0, 0.02, 0.03, 0.01)
        S.fit_model(frame, Pd, Qd, bud, bid, 3.0, learning_rate=0.02, embedding_regularization=0.03, bias_regularization=0.01)
    S.check_fit(frame)
    for got, want, key in ((Pd, P, "P"), (Qd, Q, "Q"), (bud, bu, "bu"), (bid, bi, "bi")):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=RTOL, atol=1e-14, err_msg=key)
    np.testing.assert_allclose(S.predict(u[:100], i[:100], Pd, Qd, bud, bid, 3.0).cpu().numpy(),
                               OS.predict(u[:100], i[:100], P, Q, bu, bi, 3.0), rtol=RTOL)


def test_recommend_ties_padding_and_batches(S):
    # exact scores with ties: the earlier item wins (strict '>' replacement in the reference's loop)
    Q = np.array([[1.0, 0.0], [2.0, 0.0], [2.0, 0.0], [0.5, 0.0], [2.0, 0.0], [-1.0, 0.0]])
    P = np.array([[1.0, 5.0], [-1.0, 0.0], [0.0, 0.0]])
    vals, ids = S.recommend_users(np.array([0, 1, 2], np.int32), *_dev(P, Q), 4)
    assert ids.cpu().numpy().tolist() == [[1, 2, 4, 0], [5, 3, 0, 1], [0, 1, 2, 3]]
    assert vals.cpu().numpy().tolist() == [[2.0, 2.0, 2.0, 1.0], [1.0, -0.5, -1.0, -2.0], [0.0, 0.0, 0.0, 0.0]]
    vals, ids = S.recommend_users(np.array([0], np.int32), *_dev(P, Q), 8)  # k > number of items: padded
    assert ids.cpu().numpy()[0].tolist() == [1, 2, 4, 0, 3, 5, -1, -1]
    assert np.isneginf(vals.cpu().numpy()[0, 6:]).all()
    rng = np.random.default_rng(1)
    P = rng.normal(size=(300, 50)); Q = rng.normal(size=(1000, 50))
    vals, ids = S.recommend_users(np.arange(300, dtype=np.int32), *_dev(P, Q), 10)
    s = P @ Q.T
    want = np.argsort(-s, axis=1, kind="stable")[:, :10]
    assert np.array_equal(ids.cpu().numpy(), want)
    np.testing.assert_allclose(vals.cpu().numpy(), np.take_along_axis(s, want, 1), rtol=1e-12)


def _ml1m_ratings():
    from binrec_b200 import synth
    u, i = synth.make_interactions()
    rng = np.random.default_rng(5)
    return u, i, rng.integers(1, 6, len(u)).astype(np.float64)


def test_full_ml1m_epochs_against_the_c_oracle(S):
    """BASELINE.json's ML-1M shape (1 000 209 ratings, 6040 x 3706, power-law head: the most popular item holds
    ~16 000 ratings), the reference's d = 50, two epochs, against the sequential C loop."""
    from oracle import svd as OS
    u, i, r = _ml1m_ratings()
    U, I, d = 6040, 3706, 50
    rng = np.random.default_rng(0)
    P = rng.random((U, d)) / d; Q = rng.random((I, d)) / d
    bu = rng.normal(0, 0.05, U); bi = rng.normal(0, 0.05, I)
    Pd, Qd, bud, bid = _dev(P, Q, bu, bi)
    mu = float(r.mean())
    frame = S.Ratings(u, i, r, num_users=U, num_items=I)
    for _ in range(2):
        OS.fit_epoch_c(u, i, r, P, Q, bu, bi, mu, 0.01, 0.0, 0.01)
        S.fit_model(frame, Pd, Qd, bud, bid, mu)
    S.check_fit(frame)
    for got, want, key in ((Pd, P, "P"), (Qd, Q, "Q"), (bud, bu, "bu"), (bid, bi, "bi")):
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-9, atol=1e-13, err_msg=key)
    mse, mae = OS.errors(u, i, r, P, Q, bu, bi, mu)
    np.testing.assert_allclose(S.mean_square_error(frame, Pd, Qd, bud, bid, mu), mse, rtol=1e-10)
    np.testing.assert_allclose(S.mean_absolute_error(frame, Pd, Qd, bud, bid, mu), mae, rtol=1e-10)


def test_result_does_not_depend_on_residency_or_run(S):
    """The tickets fix the order of every pair of conflicting ratings, so how many warps are resident (how far the
    device runs ahead of the file order) must not change a single bit -- and neither may a second run."""
    u, i, r = _ml1m_ratings()
    u, i, r = u[:200000], i[:200000], r[:200000]
    frame = S.Ratings(u, i, r, num_users=6040, num_items=3706)
    outs = []
    for warps in (0, 8, 1, 0):
        P, Q, bu, bi = S.init_parameters(6040, 3706, 50, seed=3)
        bu += 0.01; bi -= 0.02
        S.fit_model(frame, P, Q, bu, bi, 3.0, warps_per_sm=warps)
        S.check_fit(frame)
        outs.append((P, Q, bu, bi))
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert torch.equal(a, b)


def _planted(n=12000, U=200, I=120, seed=2):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, U, n).astype(np.int64) + 1000; i = rng.integers(0, I, n).astype(np.int64) + 50000
    return u, i, ((u + i) % 2).astype(np.float64)


def test_zero_biases_stay_zero_and_train_and_evaluate_runs(S):
    """The reference starts from zero biases and multiplies the error by the bias itself (SVD.py:205-206): they never
    move.  train_and_evaluate end to end over a chunked dataset, with the reference's call sequence."""
    from oracle import svd as OS
    u, i, r = _planted()
    ds = S.RatingChunks.split(u, i, r, 5)
    ds.use_no_test_set()
    user_ids, item_ids, uid_max, iid_max, mu = S.digest(ds)
    uv, iv, du, di, mu_ref = OS.digest(u, i, r)
    assert list(user_ids) == uv.tolist() and list(item_ids) == iv.tolist()
    np.testing.assert_allclose(mu, mu_ref, rtol=1e-13)
    assert ds.next_cross_validation_distribution() and ds.test_set_index == 4
    res = S.train_and_evaluate(ds, user_ids, item_ids, uid_max, iid_max, mu, None, epochs=3, seed=0)
    P, Q, bu, bi = res["parameters"]
    assert not bu.any().item() and not bi.any().item()
    assert res["epoch_mse"][0] > res["epoch_mse"][-1] > 0 and len(res["epoch_test_mse"]) == 3
    assert set(res["all_data"]) == {"tp", "tn", "fp", "fn", "precision", "recall", "hitRate"}
    assert res["all_data"]["tp"] + res["all_data"]["fp"] == 10 * (uid_max + 1)
    # the same three epochs through the oracle, from the same start, over the same training chunks in order
    P0, Q0, bu0, bi0 = S.init_parameters(uid_max + 1, iid_max + 1, S.NUMBER_OF_EMBEDDINGS, seed=0)
    Pc, Qc, buc, bic = (t.cpu().numpy().copy() for t in (P0, Q0, bu0, bi0))
    n_train = sum(len(c[0]) for c in ds)
    for _ in range(3):
        OS.fit_epoch_c(du[:n_train], di[:n_train], r[:n_train], Pc, Qc, buc, bic, mu, S.LEARNING_RATE,
                       S.EMBEDDING_REGULARIZATION, S.BIAS_REGULARIZATION)
    np.testing.assert_allclose(P.cpu().numpy(), Pc, rtol=RTOL, atol=1e-14)
    np.testing.assert_allclose(Q.cpu().numpy(), Qc, rtol=RTOL, atol=1e-14)
    mse_test, _ = OS.errors(du[n_train:], di[n_train:], r[n_train:], Pc, Qc, buc, bic, mu)
    np.testing.assert_allclose(res["mse"], mse_test, rtol=RTOL)


def test_cross_validate_is_the_reference_main_loop(S):
    """Five folds, last chunk held out first, metric dicts averaged with getAverage (SVD.py:540-566); the grundfos
    variant keeps only rating == 1 rows in the test set (:389-392)."""
    u, i, r = _planted(n=6000, U=80, I=60, seed=4)
    ds = S.RatingChunks.split(u, i, r, 5, test_positive_only=True)
    out = S.cross_validate(ds, epochs=1, seed=1)
    assert out["folds"] == [4, 3, 2, 1, 0] and len(out["mse"]) == 5
    assert set(out["all_data"]) == {"tp", "tn", "fp", "fn", "precision", "recall", "hitRate"}
    ds.test_set_index = 2
    tu, ti, tr = ds.get_test_set()
    assert (tr == 1).all() and len(ds.test_frame()) == len(tu) < len(ds.chunks[2][0])
    assert len(ds.train_frame()) == sum(len(ds.chunks[c][0]) for c in (0, 1, 3, 4))
