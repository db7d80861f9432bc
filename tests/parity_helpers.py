"""Shared checks of the parity tests (test infrastructure)."""
import itertools

import numpy as np

TIE_RTOL = 1e-10  # two logL1 values closer than this (relative) are a rounding-level tie
# Brent h2, engine vs oracle, absolute.  Optim's Brent stops at |x - mid| <= 2 tol - (hi - lo)/2 with
# tol = sqrt(eps)|x| + eps, so the reference's own h2 is defined to a few 1e-8 only: measured against the exact optimum
# (50-digit arithmetic, profiles/arbiter_r02.json) the oracle's Brent is off by up to 3.2e-8, the engine's by up to
# 2.5e-8, and they differ from each other by up to 3.1e-8.  north_star's 1e-8 is not attainable through Brent by any
# two FP64 implementations; 2e-7 is the bound asserted.
H2_TOL = 2e-7


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def replay_tmax(prof, grid, flips=()):
    """tmax! (src/bulkscan_helpers.jl:330-350) on one entry's logL1 profile; comparisons whose index is in `flips`
    take the other branch.  Returns (h2 value, [(k, |margin|)])."""
    mx, cnt, margins = prof[0], 0, []
    for k in range(1, len(prof)):
        better = mx < prof[k]
        margins.append((k, abs(prof[k] - mx)))
        if k in flips:
            better = not better
        if better:
            mx = max(mx, prof[k])
            cnt += 1
    return grid[min(cnt, len(grid) - 1)], margins


def assert_h2_panel_explained(engine_panel, ref_panel, profile, grid, mode="reference"):
    """Every entry where the engine's h2_panel differs from the oracle's must be explained by a rounding-level tie:
    flipping only comparisons whose operands agree to TIE_RTOL reproduces the engine's value.  No blanket allowance.
    Returns the number of such entries."""
    grid = np.asarray(grid, dtype=np.float64)
    prof = np.stack(profile)  # [K, p, m]
    mism = np.argwhere(engine_panel != ref_panel)
    for i, j in mism:
        pr = prof[:, i, j]
        scale = max(1.0, float(np.max(np.abs(pr))))
        if mode == "argmax":
            k_eng = int(np.argmin(np.abs(grid - engine_panel[i, j])))
            assert engine_panel[i, j] == grid[k_eng]
            assert abs(pr[k_eng] - pr.max()) <= TIE_RTOL * scale, (i, j, pr, engine_panel[i, j])
            continue
        _, margins = replay_tmax(pr, grid)
        ties = [k for k, mg in margins if mg <= TIE_RTOL * scale]
        assert ties, f"h2_panel[{i},{j}]: engine {engine_panel[i, j]} vs oracle {ref_panel[i, j]} without any tie: {pr}"
        ok = False
        for r in range(1, min(len(ties), 4) + 1):
            for fl in itertools.combinations(ties, r):
                if replay_tmax(pr, grid, flips=set(fl))[0] == engine_panel[i, j]:
                    ok = True
                    break
            if ok:
                break
        assert ok, f"h2_panel[{i},{j}]: engine value {engine_panel[i, j]} is not a tie resolution of {pr}"
    return len(mism)
