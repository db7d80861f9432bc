"""The oracle against every known-answer test of the reference that needs no data file, and against
the one data fixture that survives in the checkout (the 79 x 79 BXD kinship, committed as
tests/golden/bxd_kinship.npy by tests/golden/make_fixtures.py)."""
import math
import os

import numpy as np
import pytest

import blmm_oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def lod2r(lod, n):
    return math.sqrt(1.0 - 10.0 ** (-2.0 * lod / n))


def test_r2lod_inverse():
    """test/bulkscan_test.jl:9-19."""
    assert abs(orc.r2lod(lod2r(3.0, 79), 79) - 3.0) < 1e-7


def test_makeweights_h2_one_throws():
    """test/lmm_test.jl:12-18."""
    with pytest.raises(orc.OracleError) as e:
        orc.make_weights(1.0, np.ones(4))
    assert e.value.msg == "Heritability of 1 is not allowed."


def test_gridbrent_kat():
    """test/gridbrent_test.jl:2-8: minimiser of -(x^3 + 0.2 (x-2)^2 + 3) on [-3, 1] with 100 intervals is 1."""
    r = orc.gridbrent(lambda x: -(x ** 3 + 0.2 * (x - 2) ** 2 + 3), -3.0, 1.0, 100)
    assert abs(r.minimizer - 1.0) < 1e-6


def test_brent_matches_scipy_on_smooth_function():
    from scipy.optimize import minimize_scalar
    f = lambda x: (x - 0.3) ** 2 + 0.1 * math.sin(5 * x)
    r = orc.brent_minimize(f, 0.0, 1.0)
    s = minimize_scalar(f, bounds=(0.0, 1.0), method="bounded", options={"xatol": 1e-12})
    assert abs(r.minimizer - s.x) < 1e-6


def test_compute_r_equals_correlation():
    """test/bulkscan_test.jl:25-54: computeR_LMM(Y, X, 1) == cor(X, Y) on 100 x 100 Gaussian data."""
    rng = np.random.default_rng(0)
    X, Y = rng.standard_normal((100, 100)), rng.standard_normal((100, 100))
    R = orc.compute_r_lmm(Y, X, np.ones((100, 1)))
    Xc = (X - X.mean(0)) / X.std(0)
    Yc = (Y - Y.mean(0)) / Y.std(0)
    ref = Xc.T @ Yc / 100
    assert np.sum((R - ref) ** 2) <= 1e-8


def test_wls_cholesky_equals_qr_and_ols():
    """test/wls_basic_test.jl:34-74, test/wls_results_test.jl:89-116."""
    rng = np.random.default_rng(1)
    n = 60
    X = np.column_stack([np.ones(n), rng.standard_normal((n, 2))])
    y = X @ np.array([1.0, 2.0, -1.0]) + rng.standard_normal(n)
    w = rng.uniform(0.5, 2.0, n)
    a = orc.wls(y[:, None], X, w, [0.0, 0.0], method="qr")
    b = orc.wls(y[:, None], X, w, [0.0, 0.0], method="cholesky")
    assert np.allclose(a.b, b.b, atol=1e-10) and abs(a.ell - b.ell) < 1e-9
    sw = np.sqrt(w)
    beta = np.linalg.lstsq(X * sw[:, None], y * sw, rcond=None)[0]
    assert np.allclose(a.b[:, 0], beta, atol=1e-10)
    assert abs(orc.rss(y[:, None], X)[0, 0] - np.sum((y - X @ np.linalg.lstsq(X, y, rcond=None)[0]) ** 2)) < 1e-8


def test_bxd_kinship_fixture_properties():
    """The reference ships the BXD kinship (test/run-lmmlite_R/processed_bxdData/BXDkinship.csv); the
    oracle's decomposition reproduces it and rotation == eigen(K).vectors' * y
    (test/transform_helpers_test.jl:42-53)."""
    K = np.load(os.path.join(GOLD, "bxd_kinship.npy"))
    assert K.shape == (79, 79) and np.allclose(K, K.T)
    assert np.allclose(np.diag(K), 1.0)
    Ut, lam = orc.decompose(K)
    assert lam.min() > 0.02 and lam.max() < 41.0  # SURVEY: lambda in [0.0215, 40.6]
    assert np.max(np.abs(Ut.T @ np.diag(lam) @ Ut - K)) < 1e-12
    rng = np.random.default_rng(2)
    y, g = rng.standard_normal((79, 3)), rng.standard_normal((79, 5))
    Y0, X0, l2 = orc.transform_rotation(y, g, K)
    assert np.allclose(Y0, Ut @ y) and np.allclose(X0[:, 1:], Ut @ g) and np.allclose(X0[:, 0], Ut @ np.ones(79))
    Us, S = orc.decompose(K, "svd")
    assert np.allclose(np.sort(S), np.sort(lam), atol=1e-10)


def test_golden_oracle_outputs_reproduce():
    """Regression pin: the committed oracle outputs on the BXD kinship + seeded synthetic traits."""
    z = np.load(os.path.join(GOLD, "oracle_small.npz"))
    K = np.load(os.path.join(GOLD, "bxd_kinship.npy"))
    grid = np.arange(10) / 10.0
    r = orc.bulkscan_null_grid(z["Y"], z["G"], K, grid)
    assert np.array_equal(r.h2_null_list, z["null_h2"])
    assert np.max(np.abs(r.L - z["null_L"])) < 1e-10
    a = orc.bulkscan_alt_grid(z["Y"], z["G"], K, grid)
    assert np.max(np.abs(a.L - z["alt_L"])) < 1e-10
    assert np.array_equal(a.h2_panel, z["alt_h2_panel"])  # same code, same machine arithmetic: bit for bit


def test_golden_single_trait_paths_reproduce():
    """tests/golden/oracle_scan.npz (make_fixtures.py): scan null / permutations / assumption="alt", bulkscan_null,
    lod2log10p and get_thresholds on the real BXD kinship."""
    z = np.load(os.path.join(GOLD, "oracle_scan.npz"))
    s = np.load(os.path.join(GOLD, "oracle_small.npz"))
    K = np.load(os.path.join(GOLD, "bxd_kinship.npy"))
    Y, G = s["Y"], s["G"]
    Ut, lam = z["Ut"], z["lam"]
    y = Y[:, 3:4]
    r = orc.scan(y, G, K, covar=z["covar"], reml=True, Ut=Ut, lam=lam)
    assert np.allclose(r["lod"], z["null_lod"], rtol=1e-10, atol=1e-12) and abs(r["h2_null"] - z["null_h2"]) < 1e-9
    rp = orc.scan(y, G, K, permutation_test=True, perm_idx=z["perm_idx"], Ut=Ut, lam=lam)
    assert np.allclose(rp["L_perms"], z["perm_L"], rtol=1e-10, atol=1e-12)
    assert np.allclose(orc.get_thresholds(rp["L_perms"], [0.1, 0.05])["thrs"], z["thr"], rtol=1e-12)
    ra = orc.scan(y, G[:, :60], K, assumption="alt", prior_variance=float(np.var(y, ddof=1)), prior_sample_size=0.1,
                  Ut=Ut, lam=lam)
    assert np.allclose(ra["lod"], z["alt_lod"], rtol=1e-9, atol=1e-10)
    assert np.allclose(ra["h2_each_marker"], z["alt_h2_each"], atol=1e-9)
    assert np.allclose(orc.lod2log10p(z["null_lod"], 1), z["log10p"], rtol=1e-12)


def test_lod2log10p_known_values():
    """src/util.jl:199-206 against textbook chi-square quantiles: the 5 % / 1 % critical values of chi2(1), chi2(2)."""
    ln10 = np.log(10.0)
    for df, crit, p in ((1, 3.841458820694124, 0.05), (1, 6.634896601021213, 0.01), (2, 5.991464547107979, 0.05),
                        (2, 9.210340371976182, 0.01)):
        assert abs(orc.lod2log10p(crit / (2 * ln10), df) - (-np.log10(p))) < 1e-12
    # far tail: log-space evaluation stays finite where log(sf) underflows
    assert np.isfinite(orc.lod2log10p(500.0, 1)) and 501 < orc.lod2log10p(500.0, 1) < 502.5
