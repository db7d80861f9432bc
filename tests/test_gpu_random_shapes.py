"""Randomised shapes and options through the C-ABI against the oracle: n (every K-chunk count and the K-streamed
fallback beyond n = 100), p and m around the tile edges, 1-4 covariate columns, grids of 1-23 points, ML / REML,
priors — for null-grid, alt-grid (both h2-panel modes) and permutations."""
import numpy as np
import pytest

from parity_helpers import assert_h2_panel_explained

import blmm_oracle as orc
from blmm_b200 import bulkscan_alt_grid, bulkscan_null_grid, scan, synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


CASES = []
_rng = np.random.default_rng(20260101)
for _i in range(14):
    CASES.append(dict(n=int(_rng.choice([9, 20, 21, 40, 41, 60, 79, 80, 100, 101, 137])), p=int(_rng.integers(1, 200)),
                      m=int(_rng.integers(1, 300)), ncov=int(_rng.integers(0, 4)), nk=int(_rng.choice([1, 2, 5, 10, 23])),
                      reml=bool(_rng.integers(0, 2)), pss=float(_rng.choice([0.0, 0.1, 1.0])), seed=1000 + _i))


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c['n']}-p{c['p']}-m{c['m']}-c{c['ncov']}-k{c['nk']}")
def test_random_grid_scans(engine, case):
    n, p, m = case["n"], case["p"], case["m"]
    Y, G, K = synth.make_problem(n, p, m, seed_g=case["seed"], seed_y=case["seed"] + 1)
    Cv = synth.make_covar(n)[:, :case["ncov"]] if case["ncov"] else None
    grid = np.sort(np.random.default_rng(case["seed"]).uniform(0.0, 0.95, size=case["nk"]))
    Ut, lam = orc.decompose(K)
    dec = (np.asfortranarray(Ut.T), lam)
    kw = dict(Covar=Cv, reml=case["reml"], prior_variance=1.0, prior_sample_size=case["pss"])
    r = bulkscan_null_grid(Y, G, K, grid, decomposition=dec, engine=engine, **kw)
    ref = orc.bulkscan_null_grid(Y, G, K, grid, Ut=Ut, lam=lam, **kw)
    assert np.array_equal(r.h2_null_list, ref.h2_null_list)
    assert rel(r.L, ref.L) < 1e-8
    a = bulkscan_alt_grid(Y, G, K, grid, decomposition=dec, engine=engine, **kw)
    prof = []
    aref = orc.bulkscan_alt_grid(Y, G, K, grid, Ut=Ut, lam=lam, profile=prof, **kw)
    assert rel(a.L, aref.L) < 1e-8
    assert_h2_panel_explained(a.h2_panel, aref.h2_panel, prof, grid)
    assert np.all(np.isin(a.h2_panel, grid))
    am = bulkscan_alt_grid(Y, G, K, grid, decomposition=dec, engine=engine, h2_panel_mode="argmax", **kw)
    assert rel(am.L, aref.L) < 1e-8 and np.all(np.isin(am.h2_panel, grid))


@pytest.mark.parametrize("n,p,nperms", [(20, 33, 5), (79, 130, 300), (101, 64, 129)])
def test_random_permutation_scans(engine, n, p, nperms):
    Y, G, K = synth.make_problem(n, p, 3, seed_g=n + p, seed_y=n)
    Ut, lam = orc.decompose(K)
    dec = (np.asfortranarray(Ut.T), lam)
    idx = orc.make_perm_indices(n, nperms, 5)
    r = scan(Y[:, 1], G, K, permutation_test=True, perm_idx=idx, decomposition=dec, engine=engine)
    ref = orc.scan(Y[:, 1], G, K, permutation_test=True, perm_idx=idx, Ut=Ut, lam=lam)
    assert abs(r.h2_null - ref["h2_null"]) < 1e-6
    assert rel(r.lod, ref["lod"]) < 2e-5 and rel(r.L_perms, ref["L_perms"]) < 2e-5
    assert rel(r.max_lod, np.max(ref["L_perms"], axis=0)) < 2e-5
