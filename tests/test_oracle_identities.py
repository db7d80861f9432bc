"""The data-independent identities the reference's own test-suite asserts between its code paths
(SURVEY section 4(3)), re-stated on synthetic data for the oracle.  The same identities are asserted
for the GPU engine in tests/test_gpu_identities.py."""
import os

import numpy as np
import pytest

import blmm_oracle as orc
from blmm_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
GRID = np.arange(10) / 10.0


@pytest.fixture(scope="module")
def prob():
    K = np.load(os.path.join(GOLD, "bxd_kinship.npy"))
    G = synth.make_geno(79, 120, seed=21)
    Y = synth.make_pheno(G, K, 12, seed=22)
    return Y, G, K


def interior_trait(Y, G, K, **kw):
    """First trait whose Brent h2 is away from the boundary (a boundary fit ties with grid point 0)."""
    for j in range(Y.shape[1]):
        s = orc.scan(Y[:, j], G, K, **kw)
        if 0.05 < s["h2_null"] < 0.85:
            return j, s
    raise AssertionError("no interior trait")


def test_null_exact_equals_scan(prob):
    """test/bulkscan_test.jl:60-80."""
    Y, G, K = prob
    r = orc.bulkscan_null(Y[:, :3], G, K, reml=True)
    for j in range(3):
        s = orc.scan(Y[:, j], G, K, reml=True)
        assert abs(r.h2_null_list[j] - s["h2_null"]) < 1e-6
        assert np.sum((r.L[:, j] - s["lod"]) ** 2) <= 1e-7


def test_null_grid_with_exact_h2_equals_scan(prob):
    """test/bulkscan_test.jl:86-107: a grid that contains the Brent h2 reproduces scan()."""
    Y, G, K = prob
    j, s = interior_trait(Y, G, K)
    grid = np.sort(np.append(GRID, s["h2_null"]))
    r = orc.bulkscan_null_grid(Y[:, j:j + 1], G, K, grid)
    assert r.h2_null_list[0] == s["h2_null"]
    assert np.sum((r.L[:, 0] - s["lod"]) ** 2) <= 1e-7


def test_alt_grid_dominates_null_grid(prob):
    """alt-grid maximises over h2 per marker, so its LOD is >= the null-grid LOD (same grid)."""
    Y, G, K = prob
    a = orc.bulkscan_alt_grid(Y, G, K, GRID)
    r = orc.bulkscan_null_grid(Y, G, K, GRID)
    assert np.all(a.L >= r.L - 1e-9)
    assert set(np.unique(a.h2_panel)).issubset(set(GRID))


def test_wrapper_equals_direct(prob):
    """test/bulkscan_test.jl:139-178."""
    Y, G, K = prob
    assert np.array_equal(orc.bulkscan(Y, G, K, method="null-grid")["L"], orc.bulkscan_null_grid(Y, G, K, GRID).L)
    assert np.array_equal(orc.bulkscan(Y, G, K, method="alt-grid")["L"], orc.bulkscan_alt_grid(Y, G, K, GRID).L)


def test_svd_equals_eigen_and_covariates(prob):
    """test/scan_covar_test.jl:10-37."""
    Y, G, K = prob
    Z = synth.make_covar(79, seed=4)
    a = orc.bulkscan_null_grid(Y, G, K, GRID, Covar=Z, decomp_scheme="eigen")
    b = orc.bulkscan_null_grid(Y, G, K, GRID, Covar=Z, decomp_scheme="svd")
    assert np.array_equal(a.h2_null_list, b.h2_null_list)
    assert np.mean(np.abs(a.L - b.L)) <= 1e-8
    j, s = interior_trait(Y, G, K, covar=Z)
    grid = np.sort(np.append(GRID, s["h2_null"]))
    r = orc.bulkscan_null_grid(Y[:, j:j + 1], G, K, grid, Covar=Z)
    assert np.mean(np.abs(r.L[:, 0] - s["lod"])) <= 1e-8


def test_weights_equal_preweighting(prob):
    """test/weighted_error_test.jl:28-142: weights= is the same as pre-scaling y, G, covar, K."""
    Y, G, K = prob
    n = 79
    w = np.random.default_rng(3).uniform(0.5, 1.5, n)
    a = orc.bulkscan_null_grid(Y, G, K, GRID, weights=w)
    C = w[:, None] * np.ones((n, 1))
    b = orc.bulkscan_null_grid(w[:, None] * Y, w[:, None] * G, w[:, None] * K * w[None, :], GRID, Covar=C,
                               addIntercept=False)
    assert np.max(np.abs(a.L - b.L)) < 1e-10
    one = orc.bulkscan_null_grid(Y, G, K, GRID, weights=np.ones(n))
    assert np.max(np.abs(one.L - orc.bulkscan_null_grid(Y, G, K, GRID).L)) < 1e-10


def test_permutations_keep_original_first(prob):
    """test/transform_helpers_test.jl:116-131 and scan(...; permutation_test=true) column split."""
    Y, G, K = prob
    perm = synth.make_perm_indices(79, 16, rndseed=1)
    r = orc.scan(Y[:, 0], G, K, permutation_test=True, perm_idx=perm)
    s = orc.scan(Y[:, 0], G, K)
    assert np.sum((r["lod"] - s["lod"]) ** 2) <= 1e-7
    assert r["L_perms"].shape == (120, 16)
    ident = np.tile(np.arange(79, dtype=np.int32)[:, None], (1, 2))
    r2 = orc.scan(Y[:, 0], G, K, permutation_test=True, perm_idx=ident)
    assert np.allclose(r2["L_perms"][:, 0], r2["lod"], atol=1e-12)
    t = orc.get_thresholds(r["L_perms"], [0.1, 0.05])
    assert np.allclose(t["thrs"], np.quantile(r["L_perms"].max(axis=0), [0.9, 0.95]))


def test_errors_match_reference_strings(prob):
    Y, G, K = prob
    with pytest.raises(orc.OracleError) as e:
        orc.transform_rotation(Y, G[:-1], K)
    assert e.value.msg == "Dimension mismatch."
    with pytest.raises(orc.OracleError) as e:
        orc.scan(Y[:, 0], G, K, addIntercept=False)
    assert e.value.msg == "Intercept has to be added when no other covariate is given."
    with pytest.raises(orc.OracleError) as e:
        orc.scan_perms_lite(Y[:, :2], G, np.ones((79, 1)), K, np.zeros((79, 1), dtype=np.int32))
    assert e.value.msg == "Can only handle one trait."
    Gm = G.copy()
    Gm[:, 3] = 1.0
    with pytest.raises(orc.OracleError) as e:
        orc.bulkscan_null_grid(Y, Gm, K, GRID)
    assert e.value.msg == "Dividing by zeros: the input vector can not contain any zeros!"
