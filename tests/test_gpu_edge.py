"""GPU edge cases, error paths and the reference's self-consistency identities through the C-ABI."""
import os

import numpy as np
import pytest

from parity_helpers import assert_h2_panel_explained, H2_TOL

import blmm_oracle as orc
from blmm_b200 import BlmmError, bulkscan, bulkscan_alt_grid, bulkscan_null_grid, scan, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
GRID = np.arange(10) / 10.0


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def make(n, p, m, seed=0):
    Y, G, K = synth.make_problem(n, p, m, seed_g=100 + seed, seed_y=200 + seed)
    Ut, lam = orc.decompose(K)
    return Y, G, K, Ut, lam, (np.asfortranarray(Ut.T), lam)


@pytest.mark.parametrize("n,p,m", [(79, 1, 1), (79, 63, 127), (79, 65, 129), (20, 40, 30), (41, 100, 17), (100, 70, 50),
                                   (8, 5, 3)])
def test_shapes_null_and_alt(engine, n, p, m):
    """Ragged sizes around the tile edges (64 markers / 128 traits) and every K-chunk count 1..5."""
    Y, G, K, Ut, lam, dec = make(n, p, m, seed=n + p)
    r = bulkscan_null_grid(Y, G, K, GRID, decomposition=dec, engine=engine)
    ref = orc.bulkscan_null_grid(Y, G, K, GRID, Ut=Ut, lam=lam)
    assert np.array_equal(r.h2_null_list, ref.h2_null_list)
    assert rel(r.L, ref.L) < 1e-8
    a = bulkscan_alt_grid(Y, G, K, GRID, decomposition=dec, engine=engine)
    aref = orc.bulkscan_alt_grid(Y, G, K, GRID, Ut=Ut, lam=lam)
    assert rel(a.L, aref.L) < 1e-8


def test_golden_bxd_kinship_fixture(engine):
    """The committed fixture on the real BXD kinship (tests/golden/make_fixtures.py), own eigendecomposition."""
    z = np.load(os.path.join(GOLD, "oracle_small.npz"))
    K = np.load(os.path.join(GOLD, "bxd_kinship.npy"))
    r = bulkscan_null_grid(z["Y"], z["G"], K, GRID, engine=engine)  # cuSOLVER syevd inside
    assert np.array_equal(r.h2_null_list, z["null_h2"])
    assert rel(r.L, z["null_L"]) < 1e-8
    a = bulkscan_alt_grid(z["Y"], z["G"], K, GRID, engine=engine)
    assert rel(a.L, z["alt_L"]) < 1e-8
    prof = []
    orc.bulkscan_alt_grid(z["Y"], z["G"], K, GRID, profile=prof)
    assert_h2_panel_explained(a.h2_panel, z["alt_h2_panel"], prof, GRID)


def test_grid_of_twenty_and_argmax_panel(engine):
    """The reference's own alt-grid test grid 0:0.05:0.95 (test/bulkscan_test.jl:113-137)."""
    Y, G, K, Ut, lam, dec = make(79, 90, 40, seed=3)
    grid = np.arange(20) / 20.0
    a = bulkscan_alt_grid(Y, G, K, grid, decomposition=dec, engine=engine)
    prof = []
    ref = orc.bulkscan_alt_grid(Y, G, K, grid, Ut=Ut, lam=lam, profile=prof)
    assert rel(a.L, ref.L) < 1e-8
    assert_h2_panel_explained(a.h2_panel, ref.h2_panel, prof, grid)
    am = bulkscan_alt_grid(Y, G, K, grid, decomposition=dec, engine=engine, h2_panel_mode="argmax")
    assert rel(am.L, ref.L) < 1e-8
    # arg-max panel: the grid value at which the alternative log-likelihood peaks
    Y0, X0, l0 = orc.transform_rotation(Y, G, K, Ut=Ut, lam=lam)
    ell = orc.grid_loglik(Y0, X0[:, :1], l0, grid, [1.0, 0.0])
    ll1 = np.stack([orc.weighted_liteqtl(Y0, X0, l0, h) * orc.LN10 + ell[k][None, :] for k, h in enumerate(grid)])
    want = grid[np.argmax(ll1, axis=0)]
    assert_h2_panel_explained(am.h2_panel, want, list(ll1), grid, mode="argmax")
    one = bulkscan_alt_grid(Y, G, K, [0.3], decomposition=dec, engine=engine)
    assert rel(one.L, orc.weighted_liteqtl(Y0, X0, l0, 0.3)) < 1e-8
    assert np.all(one.h2_panel == 0.3)


def test_identities_through_the_engine(engine):
    """null-grid with the exact h2 in the grid == the per-marker QR scan (test/bulkscan_test.jl:86-107);
    weights == pre-weighting (test/weighted_error_test.jl); svd == eigen (test/scan_covar_test.jl)."""
    Y, G, K, Ut, lam, dec = make(79, 80, 10, seed=7)
    for j in range(10):
        s = orc.scan(Y[:, j], G, K, Ut=Ut, lam=lam)
        if 0.05 < s["h2_null"] < 0.85:
            break
    grid = np.sort(np.append(GRID, s["h2_null"]))
    r = bulkscan_null_grid(Y[:, j:j + 1], G, K, grid, decomposition=dec, engine=engine)
    assert r.h2_null_list[0] == s["h2_null"]
    assert np.sum((r.L[:, 0] - s["lod"]) ** 2) <= 1e-7
    w = np.random.default_rng(3).uniform(0.5, 1.5, 79)
    a = bulkscan_null_grid(Y, G, K, GRID, weights=w, engine=engine)
    b = orc.bulkscan_null_grid(Y, G, K, GRID, weights=w)
    assert rel(a.L, b.L) < 1e-8
    e = bulkscan_null_grid(Y, G, K, GRID, engine=engine, decomp_scheme="eigen")
    v = bulkscan_null_grid(Y, G, K, GRID, engine=engine, decomp_scheme="svd")
    assert np.array_equal(e.h2_null_list, v.h2_null_list) and np.mean(np.abs(e.L - v.L)) <= 1e-8
    wrap = bulkscan(Y, G, K, engine=engine)  # defaults: null-grid, 0:0.1:0.9
    assert np.array_equal(wrap.L, e.L)
    pv = bulkscan(Y, G, K, engine=engine, output_pvals=True)
    assert rel(pv.log10Pvals_mat, orc.lod2log10p(e.L, 1)) < 1e-8


def test_scan_perms_given_decomposition_exact(engine):
    """Permutation LODs at 1e-8 once the h2 estimate is taken out of the comparison: with a one-point
    Brent bracket both sides land on the same h2 to ~1e-9, and LODs follow."""
    Y, G, K, Ut, lam, dec = make(79, 150, 4, seed=9)
    perm = synth.make_perm_indices(79, 200, rndseed=3)
    r = scan(Y[:, 1], G, K, permutation_test=True, perm_idx=perm, decomposition=dec, engine=engine)
    ref = orc.scan(Y[:, 1], G, K, permutation_test=True, perm_idx=perm, Ut=Ut, lam=lam)
    # oracle LODs recomputed at the ENGINE's h2 (removes the Brent 1e-8 wobble, SURVEY hard part 3a)
    y0, X0, l0 = orc.transform_rotation(Y[:, 1:2], G, K, Ut=Ut, lam=lam)
    w = orc.make_weights(r.h2_null, l0)
    est = orc.wls(y0, X0[:, :1], w, [0.0, 0.0])
    r0 = (y0 - X0[:, :1] @ est.b) * np.sqrt(w)[:, None]
    X00 = orc.resid(X0[:, 1:] * np.sqrt(w)[:, None], X0[:, :1] * np.sqrt(w)[:, None])
    rp = orc.shuffle_vector(r0[:, 0], perm)
    rp = rp / np.linalg.norm(rp, axis=0)
    X00 = X00 / np.linalg.norm(X00, axis=0)
    Lref = orc.r2lod(X00.T @ rp, 79)
    assert rel(r.lod, Lref[:, 0]) < 1e-8
    assert rel(r.L_perms, Lref[:, 1:]) < 1e-8
    assert np.array_equal(np.argmax(r.L_perms, axis=0), np.argmax(Lref[:, 1:], axis=0))
    assert abs(r.h2_null - ref["h2_null"]) < H2_TOL


def test_error_paths(engine):
    Y, G, K, Ut, lam, dec = make(79, 40, 6, seed=11)
    with pytest.raises(BlmmError) as e:
        bulkscan_null_grid(Y, G[:-1], K, GRID, engine=engine)
    assert e.value.msg == "Dimension mismatch."
    with pytest.raises(BlmmError) as e:
        bulkscan_null_grid(Y, G, K, [0.0, 0.5, 1.0], decomposition=dec, engine=engine)
    assert e.value.msg == "Heritability of 1 is not allowed."
    Gm = G.copy()
    Gm[:, 3] = 1.0  # monomorphic marker: the reference's colDivide! throws
    with pytest.raises(BlmmError) as e:
        bulkscan_null_grid(Y, Gm, K, GRID, decomposition=dec, engine=engine)
    assert e.value.msg == "Dividing by zeros: the input vector can not contain any zeros!"
    with pytest.raises(BlmmError) as e:
        scan(Y[:, :2], G, K, permutation_test=True, nperms=4, decomposition=dec, engine=engine)
    assert e.value.msg == "Can only handle one trait."
    with pytest.raises(BlmmError) as e:
        scan(Y[:, 0], G, K, addIntercept=False, permutation_test=True, nperms=4, engine=engine)
    assert e.value.msg == "Intercept has to be added when no other covariate is given."
    # the context stays usable after an error
    r = bulkscan_null_grid(Y, G, K, GRID, decomposition=dec, engine=engine)
    assert rel(r.L, orc.bulkscan_null_grid(Y, G, K, GRID, Ut=Ut, lam=lam).L) < 1e-8


def test_full_size_properties(engine):
    """BASELINE.json's full BXD shape through size-independent properties: alt-grid LOD >= null-grid LOD
    on the same grid, both equal the oracle on a random sample of trait columns, finite everywhere."""
    n, p, m = synth.BXD_N, synth.BXD_P, synth.BXD_M
    Y, G, K = synth.make_problem(n, p, m)
    U, lam, _ = engine.decompose(K)
    dec = (U, lam)
    a = bulkscan_alt_grid(Y, G, K, GRID, decomposition=dec, engine=engine)
    r = bulkscan_null_grid(Y, G, K, GRID, decomposition=dec, engine=engine)
    assert a.L.shape == (p, m) and np.isfinite(a.L).all() and np.isfinite(r.L).all()
    assert np.all(a.L >= r.L - 1e-9)
    assert set(np.unique(a.h2_panel)).issubset(set(GRID)) and set(np.unique(r.h2_null_list)).issubset(set(GRID))
    cols = np.random.default_rng(0).choice(m, size=48, replace=False)
    Ut = np.ascontiguousarray(U.T)
    ref = orc.bulkscan_alt_grid(Y[:, cols], G, K, GRID, Ut=Ut, lam=lam)
    assert rel(a.L[:, cols], ref.L) < 1e-8
    assert np.array_equal(np.argmax(a.L[:, cols], axis=0), np.argmax(ref.L, axis=0))
    ref0 = orc.bulkscan_null_grid(Y[:, cols], G, K, GRID, Ut=Ut, lam=lam)
    assert np.array_equal(r.h2_null_list[cols], ref0.h2_null_list)
    assert rel(r.L[:, cols], ref0.L) < 1e-8


@pytest.mark.parametrize("transfer", ["index", "f64"])
def test_alt_grid_host_chunked_copyback(engine, monkeypatch, transfer):
    """m >= 2048 traits through HOST buffers: the alt-grid scan runs in 8 trait-tile chunks whose columns
    are copied back on a second stream while the next chunk is scanned — results must be identical to the
    oracle (and to the same call with a padded leading dimension)."""
    monkeypatch.setenv("BLMM_B200_H2_TRANSFER", transfer)  # default: index only for panels of >= 1e8 entries
    monkeypatch.setenv("BLMM_B200_HOST_THREADS", "3")
    Y, G, K, Ut, lam, dec = make(79, 70, 2101, seed=31)
    a = bulkscan_alt_grid(Y, G, K, GRID, decomposition=dec, engine=engine)
    prof = []
    ref = orc.bulkscan_alt_grid(Y, G, K, GRID, Ut=Ut, lam=lam, profile=prof)
    assert rel(a.L, ref.L) < 1e-8
    assert_h2_panel_explained(a.h2_panel, ref.h2_panel, prof, GRID)
    assert np.all(np.isin(a.h2_panel, GRID))
    # The host path brings the h2 panel back as one-byte grid indices and expands them with host threads; the
    # device-resident path stores Float64 grid values from the kernel.  Same call, both ways: identical bits.
    import torch
    from blmm_b200 import _lib as L
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)).to(dev)
    n, p, m = 79, 70, 2101
    for mode, name in ((L.H2PANEL_REFERENCE, "reference"), (L.H2PANEL_ARGMAX, "argmax")):
        host = bulkscan_alt_grid(Y, G, K, GRID, decomposition=dec, engine=engine, h2_panel_mode=name)
        d = [t(Y), t(G), t(np.ones((n, 1))), t(dec[0]), torch.from_numpy(lam.copy()).to(dev)]
        dL = torch.empty((m, p), dtype=torch.float64, device=dev)
        dH = torch.empty((m, p), dtype=torch.float64, device=dev)
        pr = engine.make_problem(n, p, m, 1, *[x.data_ptr() for x in d])
        o, keep = engine.make_opts(method=L.METHOD_ALT_GRID, h2_grid=GRID, mem_space=L.MEM_DEVICE, h2_panel_mode=mode)
        engine.bulkscan_raw(pr, o, dL.data_ptr(), dH.data_ptr())
        engine.sync()
        assert np.array_equal(host.L, dL.cpu().numpy().T)
        assert np.array_equal(host.h2_panel, dH.cpu().numpy().T)


def test_ingest_straight_to_device(engine, tmp_path):
    """readBXDpheno / readGenoProb_ExcludeComplements with device=0: the parsed matrices land in device memory and feed
    a device-resident scan without ever being host arrays on the caller's side; same LODs as the host path."""
    import torch
    from blmm_b200 import _lib as L, readBXDpheno, readGenoProb_ExcludeComplements
    from test_ingest_cpu import write_geno, write_pheno
    ppath, gpath = str(tmp_path / "pheno.csv"), str(tmp_path / "geno.csv")
    Y = write_pheno(ppath, n=79, m=140, seed=5)
    G = write_geno(gpath, n=79, p=90, seed=6)
    K = synth.calc_kinship_host(G)
    dY, dG = readBXDpheno(ppath, device=0), readGenoProb_ExcludeComplements(gpath, device=0)
    assert dY.shape == Y.shape and dG.shape == G.shape
    U, lam, _ = engine.decompose(K)
    dev = torch.device("cuda:0")
    dC = torch.ones(79, dtype=torch.float64, device=dev)
    dU = torch.from_numpy(np.ascontiguousarray(U.T)).to(dev)
    dl = torch.from_numpy(lam.copy()).to(dev)
    dL = torch.empty((140, 90), dtype=torch.float64, device=dev)
    dh = torch.empty(140, dtype=torch.float64, device=dev)
    pr = engine.make_problem(79, 90, 140, 1, dY.ptr, dG.ptr, dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    o, keep = engine.make_opts(method=L.METHOD_NULL_GRID, h2_grid=GRID, mem_space=L.MEM_DEVICE)
    engine.bulkscan_raw(pr, o, dL.data_ptr(), dh.data_ptr())
    engine.sync()
    host = bulkscan_null_grid(readBXDpheno(ppath), readGenoProb_ExcludeComplements(gpath), K, GRID,
                              decomposition=(U, lam), engine=engine)
    assert np.array_equal(host.L, dL.cpu().numpy().T) and np.array_equal(host.h2_null_list, dh.cpu().numpy())
    dY.free()
    dG.free()
