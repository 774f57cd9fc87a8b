"""The biased-SVD oracle (oracle/svd.py, oracle/svd_c.c) against golden vectors produced by EXECUTING the reference's
own code (tests/golden/make_svd_golden.py -> svd_golden.npz): digest, fit_model, predict, mean errors, quintile ratings.
float64 throughout; the only freedom is the summation order inside np.dot, hence rtol 1e-12."""
import os

import numpy as np
import pytest

from oracle import svd as OS

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "svd_golden.npz"))
CASES = ["tiny", "defaults", "stars_reg", "quintiles"]


def _case(name):
    g = lambda k: G[f"{name}/{k}"]
    lr, ereg, breg, epochs, d = g("hyper")
    return g, float(lr), float(ereg), float(breg), int(epochs), int(d)


@pytest.mark.parametrize("name", CASES)
def test_digest_matches_the_reference(name):
    g, *_ = _case(name)
    uv, iv, u, i, mu = OS.digest(g("raw_users"), g("raw_items"), g("ratings"))
    assert np.array_equal(uv, g("user_vocab")) and np.array_equal(iv, g("item_vocab"))
    assert np.array_equal(u, g("users")) and np.array_equal(i, g("items"))
    np.testing.assert_allclose(mu, g("global_bias"), rtol=1e-13)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("impl", ["python", "c"])
def test_fit_predict_errors_match_the_reference(name, impl):
    g, lr, ereg, breg, epochs, d = _case(name)
    P, Q, bu, bi = g("P0").copy(), g("Q0").copy(), g("bu0").copy(), g("bi0").copy()
    mu = float(g("global_bias"))
    fit = OS.fit_epoch if impl == "python" else OS.fit_epoch_c
    for e in range(epochs):
        fit(g("users"), g("items"), g("ratings"), P, Q, bu, bi, mu, lr, ereg, breg)
        mse, mae = OS.errors(g("users"), g("items"), g("ratings"), P, Q, bu, bi, mu)
        np.testing.assert_allclose(mse, g("mse")[e], rtol=1e-11)
        np.testing.assert_allclose(mae, g("mae")[e], rtol=1e-11)
    for got, key in ((P, "P1"), (Q, "Q1"), (bu, "bu1"), (bi, "bi1")):
        np.testing.assert_allclose(got, g(key), rtol=1e-11, atol=1e-14, err_msg=key)
    pred0 = OS.predict(g("users")[:1], g("items")[:1], P, Q, bu, bi, mu)[0]
    np.testing.assert_allclose(pred0, g("pred0"), rtol=1e-12)
    # recommend(): the three best items of user 0 by plain dot product
    s = Q @ P[0]
    top = np.argsort(-s, kind="stable")[:3]
    assert top.tolist() == g("recommend_u0")[:, 1].astype(int).tolist()
    np.testing.assert_allclose(s[top], g("recommend_u0")[:, 0], rtol=1e-11)


def test_zero_biases_stay_zero_like_the_reference():
    g, lr, ereg, breg, epochs, d = _case("defaults")                       # the reference's own start: zero biases
    assert not g("bu1").any() and not g("bi1").any()


def test_quintile_ratings():
    for v, a, b in G["quintile_cases"]:
        assert OS.place_in_quintile(v, (1, 2, 4)) == a and OS.place_in_quintile(v, (1, 1, 2)) == b
    g, *_ = _case("quintiles")
    assert np.array_equal(OS.quintile_rating(g("transaction_count"), g("quantity_sum")), g("ratings"))


def test_dependency_levels_reproduce_the_sequential_pass():
    """Processing the ratings level by level (any order inside a level) gives the sequential result bit for bit."""
    g, lr, ereg, breg, epochs, d = _case("stars_reg")
    u, i, r = g("users"), g("items"), g("ratings")
    lev = OS.dependency_levels(u, i, int(u.max()) + 1, int(i.max()) + 1)
    for k in range(1, len(u)):                                             # definition: conflicts are strictly ordered
        prev = [j for j in range(k) if u[j] == u[k] or i[j] == i[k]]
        assert lev[k] == 1 + max((lev[j] for j in prev), default=0)
    P, Q, bu, bi = g("P0").copy(), g("Q0").copy(), g("bu0").copy(), g("bi0").copy()
    P2, Q2, bu2, bi2 = P.copy(), Q.copy(), bu.copy(), bi.copy()
    mu = float(g("global_bias"))
    OS.fit_epoch(u, i, r, P, Q, bu, bi, mu, lr, ereg, breg)
    rng = np.random.default_rng(0)
    for l in range(1, int(lev.max()) + 1):
        idx = rng.permutation(np.nonzero(lev == l)[0])                     # shuffled inside the level
        assert len(set(u[idx])) == len(idx) and len(set(i[idx])) == len(idx)
        OS.fit_epoch(u[idx], i[idx], r[idx], P2, Q2, bu2, bi2, mu, lr, ereg, breg)
    assert np.array_equal(P, P2) and np.array_equal(Q, Q2) and np.array_equal(bu, bu2) and np.array_equal(bi, bi2)
