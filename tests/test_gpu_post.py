"""GPU parity of the post-processing entry points against the oracle:
lod2log10p (src/util.jl:199-206), get_thresholds (src/analysis_helpers/single_trait_analysis.jl:13-23),
observation weights crossing the ABI (src/bulkscan.jl:231-250, src/scan.jl:204-222)."""
import numpy as np
import pytest

import blmm_oracle as orc
from blmm_b200 import (bulkscan, bulkscan_alt_grid, bulkscan_null, bulkscan_null_grid, get_thresholds, lod2log10p, scan,
                       synth, thresholds_from_max)

pytestmark = pytest.mark.gpu
GRID = np.arange(10) / 10.0


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


@pytest.mark.parametrize("df", [1, 2, 3, 4, 7])
def test_lod2log10p_matches_chisq_logccdf(engine, df):
    """From LOD 0 (p = 1) through tiny LODs to LOD 500 (p ~ 1e-500: the log-space form must not underflow)."""
    lod = np.concatenate([[0.0, 1e-300, 1e-12, 1e-6, 1e-3], np.linspace(0.01, 2.0, 300), np.linspace(2.0, 60.0, 300),
                          [100.0, 250.0, 500.0]])
    got = lod2log10p(lod, df, engine=engine)
    ref = orc.lod2log10p(lod, df)
    assert got.shape == lod.shape
    assert np.all(np.isfinite(got))
    # relative to max(1, |ref|) as everywhere; and tight relative accuracy where the value is tiny
    assert rel(got, ref) < 1e-10
    big = ref > 1e-6
    assert np.max(np.abs(got[big] - ref[big]) / ref[big]) < 1e-9


def test_lod2log10p_matrix_and_kat(engine):
    """test/scan_covar_test.jl:39-50 compares log10p output with lod2log10p.(L, 1); KAT: LOD of the 5 % chi-square(1)
    critical value 3.841458820694124 is -log10(0.05)."""
    rng = np.random.default_rng(5)
    Lm = rng.gamma(1.0, 1.0, size=(37, 23))
    got = lod2log10p(Lm, 1, engine=engine)
    assert got.shape == Lm.shape and rel(got, orc.lod2log10p(Lm, 1)) < 1e-10
    crit = 3.841458820694124 / (2 * np.log(10.0))
    assert abs(lod2log10p(np.array([crit]), 1, engine=engine)[0] - (-np.log10(0.05))) < 1e-12


def test_thresholds_type7_quantiles(engine):
    rng = np.random.default_rng(11)
    for nperms in (1, 2, 7, 1000, 10000):
        mx = rng.gamma(2.0, 1.0, size=nperms)
        sl = [0.1, 0.05, 0.01, 1.0, 0.0]
        got = thresholds_from_max(mx, sl, engine=engine)
        assert np.array_equal(got.thrs, orc.quantile_type7(mx, 1.0 - np.asarray(sl)))  # bit for bit
        assert np.allclose(got.thrs, np.quantile(mx, 1.0 - np.asarray(sl)), rtol=1e-14, atol=0)
    Lp = rng.gamma(1.0, 1.0, size=(50, 300))
    t = get_thresholds(Lp, [0.1, 0.05], engine=engine)
    tr = orc.get_thresholds(Lp, [0.1, 0.05])
    assert np.array_equal(t.thrs, tr["thrs"]) and np.array_equal(t.probs, tr["probs"])


def test_observation_weights_all_methods(engine):
    """weights= is applied on the device (row scaling folded into the rotation); the oracle pre-scales on the host
    exactly as the reference does.  Own decomposition of W*K*W on both sides, so LODs (sign invariant) agree."""
    Y, G, K = synth.make_problem(79, 150, 60, seed_g=21, seed_y=22)
    rng = np.random.default_rng(3)
    w = rng.uniform(0.5, 2.0, size=79)
    Cv = synth.make_covar(79)[:, :2]
    a = bulkscan_null_grid(Y, G, K, GRID, Covar=Cv, weights=w, engine=engine)
    b = orc.bulkscan_null_grid(Y, G, K, GRID, Covar=Cv, weights=w)
    assert np.array_equal(a.h2_null_list, b.h2_null_list) and rel(a.L, b.L) < 1e-8
    a = bulkscan_alt_grid(Y, G, K, GRID, Covar=Cv, weights=w, engine=engine)
    b = orc.bulkscan_alt_grid(Y, G, K, GRID, Covar=Cv, weights=w)
    assert rel(a.L, b.L) < 1e-8
    a = bulkscan_null(Y[:, :12], G, K, Covar=Cv, weights=w, reml=True, engine=engine)
    b = orc.bulkscan_null(Y[:, :12], G, K, Covar=Cv, weights=w, reml=True)
    assert np.max(np.abs(a.h2_null_list - b.h2_null_list)) < 1e-6
    assert rel(a.L, b.L) < 2e-5  # LOD sensitivity to the 1e-6 Brent tolerance
    s = scan(Y[:, 3], G, K, covar=Cv, weights=w, engine=engine)
    r = orc.scan(Y[:, 3], G, K, covar=Cv, weights=w)
    assert abs(s.h2_null - r["h2_null"]) < 1e-6 and rel(s.lod, r["lod"]) < 2e-5


def test_weights_equal_prescaling_through_engine(engine):
    """test/weighted_error_test.jl: scanning with weights == scanning pre-weighted inputs with K_st, no intercept added."""
    Y, G, K = synth.make_problem(79, 90, 33, seed_g=31, seed_y=32)
    w = np.random.default_rng(9).uniform(0.7, 1.4, size=79)
    a = bulkscan_alt_grid(Y, G, K, GRID, weights=w, engine=engine)
    Yw, Gw = w[:, None] * Y, w[:, None] * G
    Kw = w[:, None] * K * w[None, :]
    b = bulkscan_alt_grid(Yw, Gw, Kw, GRID, Covar=w[:, None], addIntercept=False, engine=engine)
    assert rel(a.L, b.L) < 1e-9


def test_output_pvals_fused(engine):
    """output_pvals through blmm_opts.chisq_df for the three methods (src/bulkscan.jl:154-160)."""
    Y, G, K = synth.make_problem(79, 130, 70, seed_g=41, seed_y=42)
    for method in ("null-grid", "alt-grid", "null-exact"):
        r = bulkscan(Y, G, K, method=method, output_pvals=True, chisq_df=1, engine=engine)
        plain = bulkscan(Y, G, K, method=method, engine=engine)
        assert np.array_equal(r.L, plain.L)
        assert r.Chisq_df == 1
        assert rel(r.log10Pvals_mat, orc.lod2log10p(np.maximum(plain.L, 0.0), 1)) < 1e-9
    r2 = bulkscan(Y, G, K, method="null-grid", output_pvals=True, chisq_df=2, engine=engine)
    assert rel(r2.log10Pvals_mat, orc.lod2log10p(np.maximum(r2.L, 0.0), 2)) < 1e-9


def test_golden_single_trait_fixture(engine):
    """The committed single-trait fixture (tests/golden/oracle_scan.npz) through the C-ABI, with the fixture's own
    decomposition and shuffle indices: scan null (REML, one covariate), permutations, assumption="alt",
    bulkscan_null, thresholds, p-values."""
    import os
    gold = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(gold, "oracle_scan.npz"))
    s = np.load(os.path.join(gold, "oracle_small.npz"))
    K = np.load(os.path.join(gold, "bxd_kinship.npy"))
    Y, G = s["Y"], s["G"]
    dec = (np.asfortranarray(z["Ut"].T), z["lam"])
    y = Y[:, 3]
    r = scan(y, G, K, covar=z["covar"], reml=True, decomposition=dec, engine=engine)
    assert abs(r.h2_null - z["null_h2"]) < 1e-6 and rel(r.lod, z["null_lod"]) < 2e-5
    rp = scan(y, G, K, permutation_test=True, perm_idx=z["perm_idx"], decomposition=dec, engine=engine)
    assert abs(rp.h2_null - z["perm_h2"]) < 1e-6
    assert rel(rp.lod, z["perm_lod"]) < 2e-5 and rel(rp.L_perms, z["perm_L"]) < 2e-5
    assert rel(get_thresholds(rp.L_perms, [0.1, 0.05], engine=engine).thrs, z["thr"]) < 2e-5
    ra = scan(y, G[:, :60], K, assumption="alt", prior_variance=float(np.var(y, ddof=1)), prior_sample_size=0.1,
              decomposition=dec, engine=engine)
    assert np.max(np.abs(ra.h2_each_marker - z["alt_h2_each"])) < 2e-6 and rel(ra.lod, z["alt_lod"]) < 1e-6
    rb = bulkscan_null(Y[:, :8], G, K, reml=True, prior_variance=0.0, decomposition=dec, engine=engine)
    assert np.max(np.abs(rb.h2_null_list - z["bnull_h2"])) < 1e-6 and rel(rb.L, z["bnull_L"]) < 2e-5
    assert rel(lod2log10p(z["null_lod"], 1, engine=engine), z["log10p"]) < 1e-10


def test_scan_pvals_and_profile_keywords(engine):
    """scan's remaining keyword surface (src/scan.jl:94-109): output_pvals / chisq_df on all three branches and
    profileLL / markerID / h2_grid (profile_LL, src/analysis_helpers/single_trait_analysis.jl:46-73)."""
    import blmm_oracle as orc
    from blmm_b200 import scan, synth
    Y, G, K = synth.make_problem(79, 90, 3, seed_g=61, seed_y=62)
    Ut, lam = orc.decompose(K)
    dec = (np.asfortranarray(Ut.T), lam)
    y = Y[:, 1]
    r = scan(y, G, K, decomposition=dec, engine=engine, output_pvals=True, chisq_df=2)
    assert np.max(np.abs(r.log10pvals - orc.lod2log10p(r.lod, 2)) / np.maximum(1.0, r.log10pvals)) < 1e-8
    a = scan(y, G, K, assumption="alt", decomposition=dec, engine=engine, output_pvals=True)
    assert np.max(np.abs(a.log10pvals - orc.lod2log10p(a.lod, 1)) / np.maximum(1.0, np.abs(a.log10pvals))) < 1e-8
    perm = synth.make_perm_indices(79, 20, 1)
    s = scan(y, G, K, permutation_test=True, perm_idx=perm, decomposition=dec, engine=engine, output_pvals=True)
    assert s.log10Pvals_perms.shape == s.L_perms.shape
    assert np.max(np.abs(s.log10pvals - orc.lod2log10p(s.lod, 1)) / np.maximum(1.0, s.log10pvals)) < 1e-8
    grid = np.arange(10) / 10.0
    res, prof = scan(y, G, K, decomposition=dec, engine=engine, reml=True, profileLL=True, markerID=7, h2_grid=grid)
    y0, X0, l0 = orc.transform_rotation(y.reshape(-1, 1), G, K, Ut=Ut, lam=lam)
    for k, h in enumerate(grid):
        w = orc.make_weights(h, l0)
        assert abs(prof.ll_list_null[k] - orc.wls(y0, X0[:, :1], w, [0.0, 0.0], reml=True).ell) < 1e-9 * 100
        Xd = np.column_stack([X0[:, 0], X0[:, 7]])  # intercept + marker 7 (1-based) = column index 7 of [1 G]
        assert abs(prof.ll_list_alt[k] - orc.wls(y0, Xd, w, [0.0, 0.0], reml=True).ell) < 1e-9 * 100
