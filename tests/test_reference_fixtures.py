"""Consumes tests/golden/ref_outputs/ — results of the REAL BulkLMM.jl on the committed inputs
(tests/golden/ref_inputs/, written by make_reference_inputs.py), produced by
`julia tests/golden/make_reference_fixtures.jl`.  When the directory exists, the CPU oracle (not-gpu tests) and the
CUDA engine (gpu tests) are checked against the reference's own numbers at BASELINE.json's tolerance; until someone
has run that script once (no Julia in the build image) these tests skip and parity stays "unpinned"."""
import os

import numpy as np
import pytest

import blmm_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
INP = os.path.join(HERE, "golden", "ref_inputs")
OUT = os.environ.get("BLMM_REF_OUTPUTS") or os.path.join(HERE, "golden", "ref_outputs")
HAVE = os.path.exists(os.path.join(OUT, "nullgrid_L.csv"))
needs_ref = pytest.mark.skipif(not HAVE, reason="tests/golden/ref_outputs absent: run julia tests/golden/make_reference_fixtures.jl")
GRID = np.arange(10) / 10.0
TOL = 1e-8
H2_TOL = 2e-6


def inp(name):
    return np.loadtxt(os.path.join(INP, name + ".csv"), delimiter=",", ndmin=2)


def out(name):
    return np.loadtxt(os.path.join(OUT, name + ".csv"), delimiter=",", ndmin=2)


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


@pytest.fixture(scope="module")
def data():
    d = dict(G=inp("G"), K=inp("K"), Y=inp("Y"), Z=inp("Covar"), w=inp("weights")[:, 0])
    if HAVE:
        d["Ut"] = np.ascontiguousarray(out("eig_U").T)
        d["lam"] = out("eig_lambda")[:, 0]
        d["dec"] = (np.asfortranarray(d["Ut"].T), d["lam"])
    return d


def test_inputs_and_scripts_are_in_place(data):
    """The pinning kit is complete without Julia: committed inputs, the generator of the inputs, the Julia script."""
    assert data["G"].shape == (79, 150) and data["Y"].shape == (79, 16) and data["K"].shape == (79, 79)
    assert np.array_equal(data["K"], data["K"].T)
    assert np.max(np.abs(np.round(orc.calc_kinship(data["G"]), 12) - data["K"])) < 2e-12
    for f in ("make_reference_inputs.py", "make_reference_fixtures.jl"):
        assert os.path.exists(os.path.join(HERE, "golden", f))
    src = open(os.path.join(HERE, "golden", "make_reference_fixtures.jl")).read()
    for call in ("bulkscan_null_grid(", "bulkscan_alt_grid(", "bulkscan_null(", "scan(", "calcKinship(", "get_thresholds(",
                 "lod2log10p.", "MersenneTwister(rndseed)"):
        assert call in src


# ---- the oracle against the reference ----------------------------------------------------------------------
@needs_ref
def test_oracle_vs_reference(data):
    d = data
    Y, G, K, Z, Ut, lam = d["Y"], d["G"], d["K"], d["Z"], d["Ut"], d["lam"]
    assert np.max(np.abs(orc.calc_kinship(G) - out("kinship"))) < 1e-13
    r = orc.bulkscan_null_grid(Y, G, K, GRID, Ut=Ut, lam=lam)
    assert np.array_equal(r.h2_null_list, out("nullgrid_h2")[:, 0]) and rel(r.L, out("nullgrid_L")) < TOL
    r = orc.bulkscan_null_grid(Y, G, K, GRID, Covar=Z, reml=True, Ut=Ut, lam=lam)
    assert np.array_equal(r.h2_null_list, out("nullgrid_cov_reml_h2")[:, 0]) and rel(r.L, out("nullgrid_cov_reml_L")) < TOL
    r = orc.bulkscan_null_grid(Y, G, K, GRID, weights=d["w"])
    assert rel(r.L, out("nullgrid_weights_L")) < TOL
    for tag, reml in (("altgrid", False), ("altgrid_reml", True)):
        a = orc.bulkscan_alt_grid(Y, G, K, GRID, reml=reml, Ut=Ut, lam=lam)
        assert rel(a.L, out(tag + "_L")) < TOL
        assert np.mean(a.h2_panel != out(tag + "_h2panel")) < 1e-3  # rounding-level ties only (LAPACK vs OpenBLAS builds)
    r = orc.bulkscan_null(Y, G, K, reml=True, prior_variance=0.0, Ut=Ut, lam=lam)
    assert np.max(np.abs(r.h2_null_list - out("nullexact_reml_h2")[:, 0])) < H2_TOL
    r2 = orc.bulkscan_null(Y, G, K, reml=True, prior_variance=0.0, Ut=Ut, lam=lam, h2_override=out("nullexact_reml_h2")[:, 0])
    assert rel(r2.L, out("nullexact_reml_L")) < TOL
    r = orc.bulkscan_null(Y, G, K, Covar=Z, optim_interval=4, Ut=Ut, lam=lam)
    assert np.max(np.abs(r.h2_null_list - out("nullexact_cov_oi4_h2")[:, 0])) < H2_TOL
    y = Y[:, 2:3]
    for tag, reml in (("ml", False), ("reml", True)):
        s = orc.scan(y, G, K, reml=reml, Ut=Ut, lam=lam)
        sc = out(f"scan_null_{tag}_scalars")[:, 0]
        assert abs(s["h2_null"] - sc[1]) < H2_TOL and abs(s["sigma2_e"] - sc[0]) < 1e-5 * sc[0]
        assert rel(s["lod"], out(f"scan_null_{tag}_lod")[:, 0]) < 1e-5
    perm = out("perms_idx0").astype(np.int64)
    s = orc.scan(y, G, K, permutation_test=True, perm_idx=perm, Ut=Ut, lam=lam)
    assert rel(s["L_perms"], out("perms_L")) < 1e-5 and rel(s["lod"], out("perms_lod")[:, 0]) < 1e-5
    assert rel(orc.get_thresholds(out("perms_L"), [0.10, 0.05])["thrs"], out("perms_thresholds")[:, 0]) < 1e-12
    lod = out("scan_null_ml_lod")[:, 0]
    assert rel(orc.lod2log10p(lod, 1), out("lod2log10p_df1")[:, 0]) < TOL
    assert rel(orc.lod2log10p(lod, 3), out("lod2log10p_df3")[:, 0]) < TOL


# ---- the engine against the reference ----------------------------------------------------------------------
@needs_ref
@pytest.mark.gpu
def test_engine_vs_reference(engine, data):
    from blmm_b200 import bulkscan_alt_grid, bulkscan_null, bulkscan_null_grid, scan, thresholds_from_max, lod2log10p
    d = data
    Y, G, K, Z, dec = d["Y"], d["G"], d["K"], d["Z"], d["dec"]
    assert np.max(np.abs(engine.calc_kinship(G) - out("kinship"))) < 1e-13
    r = bulkscan_null_grid(Y, G, K, GRID, decomposition=dec, engine=engine)
    assert np.array_equal(r.h2_null_list, out("nullgrid_h2")[:, 0]) and rel(r.L, out("nullgrid_L")) < TOL
    assert np.array_equal(np.argmax(r.L, axis=0), np.argmax(out("nullgrid_L"), axis=0))
    r = bulkscan_null_grid(Y, G, K, GRID, Covar=Z, reml=True, decomposition=dec, engine=engine)
    assert np.array_equal(r.h2_null_list, out("nullgrid_cov_reml_h2")[:, 0]) and rel(r.L, out("nullgrid_cov_reml_L")) < TOL
    r = bulkscan_null_grid(Y, G, K, GRID, weights=d["w"], engine=engine)
    assert rel(r.L, out("nullgrid_weights_L")) < TOL
    for tag, reml in (("altgrid", False), ("altgrid_reml", True)):
        a = bulkscan_alt_grid(Y, G, K, GRID, reml=reml, decomposition=dec, engine=engine)
        assert rel(a.L, out(tag + "_L")) < TOL
        assert np.array_equal(np.argmax(a.L, axis=0), np.argmax(out(tag + "_L"), axis=0))
        assert np.mean(a.h2_panel != out(tag + "_h2panel")) < 1e-3
    r = bulkscan_null(Y, G, K, reml=True, prior_variance=0.0, decomposition=dec, engine=engine)
    assert np.max(np.abs(r.h2_null_list - out("nullexact_reml_h2")[:, 0])) < H2_TOL
    assert rel(r.L, out("nullexact_reml_L")) < 1e-5  # through the Brent wobble; 1e-8 at equal h2 is test_gpu_exact.py
    y = Y[:, 2:3]
    perm = out("perms_idx0").astype(np.int32)
    s = scan(y, G, K, permutation_test=True, perm_idx=perm, decomposition=dec, engine=engine)
    sc = out("perms_scalars")[:, 0]
    assert abs(s.h2_null - sc[1]) < H2_TOL and abs(s.sigma2_e - sc[0]) < 1e-5 * sc[0]
    assert rel(s.L_perms, out("perms_L")) < 1e-5 and rel(s.lod, out("perms_lod")[:, 0]) < 1e-5
    assert np.array_equal(np.argmax(s.L_perms, axis=0), np.argmax(out("perms_L"), axis=0))
    t = thresholds_from_max(out("perms_L").max(axis=0), [0.10, 0.05], engine=engine)
    assert rel(t.thrs, out("perms_thresholds")[:, 0]) < 1e-12
    lod = out("scan_null_ml_lod")[:, 0]
    assert rel(lod2log10p(lod, 1, engine=engine), out("lod2log10p_df1")[:, 0]) < TOL
    a = scan(y, G, K, assumption="alt", decomposition=dec, engine=engine)
    assert rel(a.lod, out("scan_alt_lod")[:, 0]) < 1e-5
