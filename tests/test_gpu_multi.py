"""Host-buffer paths real callers use (pageable arrays through the pinned ring), the argument checks ADVICE asked
for, and the multi-GPU context (blmm_create_multi): one C-ABI call drives several GPUs and returns results that are
bit-identical to the one-GPU call.  The multi-GPU tests need >= 2 GPUs (gpurun --gpus 2) and skip otherwise."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GRID = np.arange(10) / 10.0


def _problem(n=79, p=600, m=2200, seed=5):
    from blmm_b200 import synth
    import blmm_oracle as orc
    Y, G, K = synth.make_problem(n, p, m, seed_g=seed, seed_y=seed + 1)
    Ut, lam = orc.decompose(K)
    return Y, G, K, np.asfortranarray(Ut.T), lam


def _raw_alt_grid(eng, Y, G, U, lam, Lbuf, Hbuf, mem_space):
    """alt-grid through the raw entry point with caller-chosen output buffers (addresses)."""
    from blmm_b200 import _lib as L
    n, m = Y.shape
    p = G.shape[1]
    keep = [np.asfortranarray(Y), np.asfortranarray(G), np.ones((n, 1)), np.asfortranarray(U), np.ascontiguousarray(lam)]
    pr = eng.make_problem(n, p, m, 1, *[a.ctypes.data for a in keep])
    o, kg = eng.make_opts(method=L.METHOD_ALT_GRID, h2_grid=GRID, mem_space=mem_space)
    eng.bulkscan_raw(pr, o, Lbuf, Hbuf)
    return keep, kg


@pytest.mark.parametrize("transfer", ["index", "f64"])
def test_pageable_results_equal_pinned_and_device(engine, transfer, monkeypatch):
    """The chunked alt-grid copy-back: pageable numpy outputs (ring + drain threads), pinned outputs (direct DMA)
    and device-resident outputs hold the same bits, for both encodings of the h2 panel over PCIe."""
    import torch
    from blmm_b200 import _lib as L
    monkeypatch.setenv("BLMM_B200_H2_TRANSFER", transfer)
    Y, G, K, U, lam = _problem()
    p, m = G.shape[1], Y.shape[1]
    # pageable
    Lp, Hp = np.full((p, m), np.nan, order="F"), np.full((p, m), np.nan, order="F")
    _raw_alt_grid(engine, Y, G, U, lam, Lp.ctypes.data, Hp.ctypes.data, L.MEM_HOST)
    # pinned
    Lq = torch.full((m, p), float("nan"), dtype=torch.float64).pin_memory()
    Hq = torch.full((m, p), float("nan"), dtype=torch.float64).pin_memory()
    _raw_alt_grid(engine, Y, G, U, lam, Lq.data_ptr(), Hq.data_ptr(), L.MEM_HOST)
    assert np.array_equal(Lp, Lq.numpy().T) and np.array_equal(Hp, Hq.numpy().T)
    assert not np.isnan(Lp).any() and not np.isnan(Hp).any()
    # device resident
    d = [torch.from_numpy(np.ascontiguousarray(a.T)).cuda() for a in (Y, G, np.ones((Y.shape[0], 1)), U)]
    dl = torch.from_numpy(lam.copy()).cuda()
    Ld = torch.empty((m, p), dtype=torch.float64, device="cuda")
    Hd = torch.empty((m, p), dtype=torch.float64, device="cuda")
    pr = engine.make_problem(Y.shape[0], p, m, 1, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(),
                             dl.data_ptr())
    o, kg = engine.make_opts(method=L.METHOD_ALT_GRID, h2_grid=GRID, mem_space=L.MEM_DEVICE)
    engine.bulkscan_raw(pr, o, Ld.data_ptr(), Hd.data_ptr())
    engine.sync()
    assert np.array_equal(Lp, Ld.cpu().numpy().T) and np.array_equal(Hp, Hd.cpu().numpy().T)


def test_pageable_large_results_other_methods(engine):
    """null-grid / null-exact / permutations with results above the ring threshold (8 MB) in pageable arrays equal
    the small-chunk (direct copy) results column for column."""
    from blmm_b200 import bulkscan, scan, synth
    Y, G, K, U, lam = _problem(p=1100, m=1000, seed=9)
    for method in ("null-grid", "null-exact"):
        big = bulkscan(Y, G, K, method=method, h2_grid=GRID, decomposition=(U, lam), engine=engine)
        small = bulkscan(Y[:, 100:200], G, K, method=method, h2_grid=GRID, decomposition=(U, lam), engine=engine)
        assert np.array_equal(big.L[:, 100:200], small.L)
    idx = synth.make_perm_indices(79, 1000, 3)
    big = scan(Y[:, 0], G, K, permutation_test=True, perm_idx=idx, decomposition=(U, lam), engine=engine)
    small = scan(Y[:, 0], G, K, permutation_test=True, perm_idx=idx[:, :50], decomposition=(U, lam), engine=engine)
    assert np.array_equal(big.L_perms[:, :50], small.L_perms)


def test_perm_idx_out_of_range_is_refused(engine):
    from blmm_b200 import BlmmError, scan, synth, _lib as L
    Y, G, K, U, lam = _problem(p=200, m=2, seed=11)
    idx = synth.make_perm_indices(79, 10, 0)
    for bad in (idx + 1, idx - 1):  # Julia's 1-based indices; a negative entry
        with pytest.raises(BlmmError) as e:
            scan(Y[:, 0], G, K, permutation_test=True, perm_idx=bad, decomposition=(U, lam), engine=engine)
        assert e.value.code == L.E_INVALID and "0-based" in e.value.msg
    ok = scan(Y[:, 0], G, K, permutation_test=True, perm_idx=idx, decomposition=(U, lam), engine=engine)
    assert np.isfinite(ok.L_perms).all()


def test_shape_checks_before_raw_pointers(engine):
    from blmm_b200 import BlmmError, bulkscan, _lib as L
    Y, G, K, U, lam = _problem(p=100, m=4, seed=12)
    for dec in ((U[:, :-1], lam), (U, lam[:-1]), (U[:-1, :-1], lam[:-1])):
        with pytest.raises(BlmmError) as e:
            bulkscan(Y, G, K, decomposition=dec, engine=engine)
        assert e.value.code == L.E_DIM and e.value.msg == "Dimension mismatch."


def test_device_mode_flags_survive_until_sync(engine):
    """A condition raised by an asynchronous device-pointer call (here a monomorphic marker: the reference throws in
    colDivide!, src/util.jl:69-71) is reported by blmm_sync even when another call was queued in between."""
    import torch
    from blmm_b200 import BlmmError, _lib as L
    Y, G, K, U, lam = _problem(p=128, m=128, seed=13)
    Gbad = G.copy()
    Gbad[:, 7] = 0.5
    n, p, m = 79, 128, 128

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()

    dY, dG, dGbad, dC, dU, dl = dev(Y), dev(G), dev(Gbad), dev(np.ones((n, 1))), dev(U), torch.from_numpy(lam.copy()).cuda()
    out = torch.empty((m, p), dtype=torch.float64, device="cuda")
    h2 = torch.empty(m, dtype=torch.float64, device="cuda")
    o, kg = engine.make_opts(method=L.METHOD_NULL_GRID, h2_grid=GRID, mem_space=L.MEM_DEVICE)
    bad = engine.make_problem(n, p, m, 1, dY.data_ptr(), dGbad.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    good = engine.make_problem(n, p, m, 1, dY.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    engine.bulkscan_raw(bad, o, out.data_ptr(), h2.data_ptr())
    engine.bulkscan_raw(good, o, out.data_ptr(), h2.data_ptr())  # used to wipe the first call's flag
    with pytest.raises(BlmmError) as e:
        engine.sync()
    assert e.value.code == L.E_ZERO_NORM
    engine.sync()  # flags are cleared once reported
    engine.bulkscan_raw(good, o, out.data_ptr(), h2.data_ptr())
    engine.sync()


# ---- several GPUs behind one context ---------------------------------------------------------------------
def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def multi():
    from blmm_b200 import Engine
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    eng = Engine(devices=list(range(min(n, 8))))
    yield eng
    eng.close()


@pytest.mark.parametrize("method", ["null-grid", "alt-grid", "null-exact"])
def test_multi_host_bulkscan_bit_equal(engine, multi, method):
    from blmm_b200 import bulkscan
    Y, G, K, U, lam = _problem(p=700, m=1003 if method != "alt-grid" else 4500, seed=21)
    kw = dict(method=method, h2_grid=GRID, decomposition=(U, lam))
    one = bulkscan(Y, G, K, engine=engine, **kw)
    many = bulkscan(Y, G, K, engine=multi, **kw)
    assert multi.device_count >= 2
    assert np.array_equal(one.L, many.L)
    if method == "alt-grid":
        assert np.array_equal(one.h2_panel, many.h2_panel)
    else:
        assert np.array_equal(one.h2_null_list, many.h2_null_list)


def test_multi_host_pvals_and_covariates(engine, multi):
    from blmm_b200 import bulkscan, synth
    Y, G, K, U, lam = _problem(p=300, m=777, seed=22)
    Cv = synth.make_covar(79)[:, :2]
    kw = dict(method="null-grid", h2_grid=GRID, decomposition=(U, lam), Covar=Cv, output_pvals=True, reml=True)
    one, many = bulkscan(Y, G, K, engine=engine, **kw), bulkscan(Y, G, K, engine=multi, **kw)
    assert np.array_equal(one.L, many.L) and np.array_equal(one.log10Pvals_mat, many.log10Pvals_mat)


def test_multi_host_perms_bit_equal(engine, multi):
    from blmm_b200 import scan, synth
    Y, G, K, U, lam = _problem(p=700, m=3, seed=23)
    for nperms in (5, 501, 2000):
        idx = synth.make_perm_indices(79, nperms, 9)
        one = scan(Y[:, 1], G, K, permutation_test=True, perm_idx=idx, decomposition=(U, lam), engine=engine)
        many = scan(Y[:, 1], G, K, permutation_test=True, perm_idx=idx, decomposition=(U, lam), engine=multi)
        assert np.array_equal(one.lod, many.lod) and np.array_equal(one.L_perms, many.L_perms)
        assert np.array_equal(one.max_lod, many.max_lod)
        assert one.h2_null == many.h2_null and one.sigma2_e == many.sigma2_e


def test_multi_errors_carry_the_reference_message(multi):
    from blmm_b200 import BlmmError, bulkscan, _lib as L
    Y, G, K, U, lam = _problem(p=300, m=600, seed=24)
    G = G.copy()
    G[:, 17] = 1.0
    with pytest.raises(BlmmError) as e:
        bulkscan(Y, G, K, decomposition=(U, lam), engine=multi)
    assert e.value.code == L.E_ZERO_NORM and e.value.msg.startswith("Dividing by zeros")


@pytest.mark.parametrize("method", ["null-grid", "alt-grid", "null-exact"])
def test_multi_device_resident_nccl_gather(engine, multi, method):
    """Device pointers on the primary GPU: NCCL scatters the trait blocks and gathers the slabs; same bits."""
    import torch
    from blmm_b200 import bulkscan, _lib as L
    Y, G, K, U, lam = _problem(p=500, m=1500, seed=25)
    n, p, m = 79, 500, 1500
    one = bulkscan(Y, G, K, method=method, h2_grid=GRID, decomposition=(U, lam), engine=engine)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a.T)).to("cuda:0")

    dY, dG, dC, dU, dl = dev(Y), dev(G), dev(np.ones((n, 1))), dev(U), torch.from_numpy(lam.copy()).to("cuda:0")
    Ld = torch.empty((m, p), dtype=torch.float64, device="cuda:0")
    alt = method == "alt-grid"
    Hd = torch.empty((m, p) if alt else (m,), dtype=torch.float64, device="cuda:0")
    pr = multi.make_problem(n, p, m, 1, dY.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    meth = {"null-grid": L.METHOD_NULL_GRID, "alt-grid": L.METHOD_ALT_GRID, "null-exact": L.METHOD_NULL_EXACT}[method]
    o, kg = multi.make_opts(method=meth, h2_grid=GRID, mem_space=L.MEM_DEVICE)
    torch.cuda.synchronize()
    for _ in range(2):
        multi.bulkscan_raw(pr, o, Ld.data_ptr(), Hd.data_ptr())
        multi.sync()
    assert multi.last_gather_ms() > 0
    assert np.array_equal(one.L, Ld.cpu().numpy().T)
    if alt:
        assert np.array_equal(one.h2_panel, Hd.cpu().numpy().T)
    else:
        assert np.array_equal(one.h2_null_list, Hd.cpu().numpy())


def test_multi_device_resident_perms_nccl_gather(engine, multi):
    import torch
    from blmm_b200 import scan, synth, _lib as L
    Y, G, K, U, lam = _problem(p=500, m=2, seed=26)
    n, p, nperms = 79, 500, 1500
    idx = synth.make_perm_indices(n, nperms, 4)
    one = scan(Y[:, 0], G, K, permutation_test=True, perm_idx=idx, decomposition=(U, lam), engine=engine)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a.T)).to("cuda:0")

    dY, dG, dC, dU, dl = dev(Y[:, :1]), dev(G), dev(np.ones((n, 1))), dev(U), torch.from_numpy(lam.copy()).to("cuda:0")
    dperm = torch.from_numpy(np.ascontiguousarray(idx.T.astype(np.int32))).to("cuda:0")
    lod = torch.empty(p, dtype=torch.float64, device="cuda:0")
    Lp = torch.empty((nperms, p), dtype=torch.float64, device="cuda:0")
    mx = torch.empty(nperms, dtype=torch.float64, device="cuda:0")
    sc = torch.empty(2, dtype=torch.float64, device="cuda:0")
    pr = multi.make_problem(n, p, 1, 1, dY.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    o, _ = multi.make_opts(prior_variance=0.0, mem_space=L.MEM_DEVICE)
    torch.cuda.synchronize()
    multi.scan_perms_raw(pr, o, dperm.data_ptr(), nperms, lod.data_ptr(), Lp.data_ptr(), mx.data_ptr(), sc.data_ptr(),
                         sc.data_ptr() + 8)
    multi.sync()
    assert np.array_equal(one.lod, lod.cpu().numpy())
    assert np.array_equal(one.L_perms, Lp.cpu().numpy().T)
    assert np.array_equal(one.max_lod, mx.cpu().numpy())
    assert multi.last_gather_ms() > 0


def test_multi_host_fit_scan_null_grid_loglik(engine, multi):
    """the other per-trait entry points of a multi-GPU context (blmm_fit_h2, blmm_scan_null, blmm_grid_loglik):
    traits sharded over the GPUs, results bit-identical to the one-GPU context"""
    Y, G, K, U, lam = _problem(p=300, m=333, seed=27)
    C = np.ones((79, 1))
    for reml in (False, True):
        a = engine.fit_h2(Y, C, U, lam, reml=reml, optim_interval=2)
        b = multi.fit_h2(Y, C, U, lam, reml=reml, optim_interval=2)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        a = engine.grid_loglik(Y, C, U, lam, GRID, reml=reml)
        b = multi.grid_loglik(Y, C, U, lam, GRID, reml=reml)
        assert np.array_equal(a, b)
    a = engine.scan_null_host(Y, G, C, U, lam, reml=True)
    b = multi.scan_null_host(Y, G, C, U, lam, reml=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # entry points that do not shard run on the primary GPU and report through the parent
    assert np.array_equal(engine.calc_kinship(G), multi.calc_kinship(G))
    assert np.array_equal(engine.lod2log10p(a[0][:, :3], 2), multi.lod2log10p(a[0][:, :3], 2))
    from blmm_b200 import BlmmError, _lib as L
    with pytest.raises(BlmmError) as e:
        multi.decompose(np.zeros((3, 4)))
    assert e.value.code == L.E_DIM


def test_multi_fewer_traits_than_gpus(engine, multi):
    """m smaller than one tile per GPU: some GPUs get no columns; permutations: fewer tiles than GPUs"""
    from blmm_b200 import bulkscan, scan, synth
    Y, G, K, U, lam = _problem(p=100, m=5, seed=28)
    for method in ("null-grid", "alt-grid", "null-exact"):
        one = bulkscan(Y, G, K, method=method, h2_grid=GRID, decomposition=(U, lam), engine=engine)
        many = bulkscan(Y, G, K, method=method, h2_grid=GRID, decomposition=(U, lam), engine=multi)
        assert np.array_equal(one.L, many.L)
    one = scan(Y[:, 0], G, K, permutation_test=True, nperms=0, perm_idx=np.zeros((79, 0), dtype=np.int32),
               decomposition=(U, lam), engine=engine)
    many = scan(Y[:, 0], G, K, permutation_test=True, nperms=0, perm_idx=np.zeros((79, 0), dtype=np.int32),
                decomposition=(U, lam), engine=multi)
    assert np.array_equal(one.lod, many.lod)


def test_contexts_come_and_go(engine):
    """contexts (single and multi-GPU) can be created, used and destroyed repeatedly, and several can coexist:
    worker threads, drain threads, streams, pinned rings and NCCL communicators are released with them"""
    from blmm_b200 import Engine, bulkscan
    Y, G, K, U, lam = _problem(p=200, m=300, seed=31)
    ref = bulkscan(Y, G, K, method="alt-grid", h2_grid=GRID, decomposition=(U, lam), engine=engine)
    devs = list(range(min(_ngpu(), 2)))
    for rep in range(3):
        e1 = Engine(0)
        e2 = Engine(devices=devs) if len(devs) > 1 else Engine(0)
        for e in (e1, e2):
            r = bulkscan(Y, G, K, method="alt-grid", h2_grid=GRID, decomposition=(U, lam), engine=e)
            assert np.array_equal(r.L, ref.L) and np.array_equal(r.h2_panel, ref.h2_panel)
        e2.close()
        e1.close()
