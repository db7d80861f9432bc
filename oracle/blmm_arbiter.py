"""Arbiter for parity disputes: the same reference formulas as blmm_oracle.py, evaluated in higher precision.

TEST INFRASTRUCTURE ONLY (like blmm_oracle.py): imported by tests/ and tests/arbiter_report.py, never by the
product.  SURVEY section 7 step 1 / section 8c ask for it: when the float64 oracle and the CUDA engine disagree near the 1e-8
tolerance, neither is "right" by construction (both round); the arbiter says which is closer to the exact value of
the reference's formula on the same float64 inputs.

Two levels:
  * `*_ld`  : numpy.longdouble (x87 80-bit, 64-bit mantissa) for whole matrices — 11 more bits than float64;
  * `*_mp`  : mpmath at 50 digits for single (marker, trait) entries and single traits — effectively exact.

The inputs are the float64 arrays both sides receive (Y, G, Covar, Ut, lambda); the rotation is part of the formula
and is redone in the higher precision.  File:line citations are to /root/reference.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

LD = np.longdouble


# ----------------------------------------------------------------------------------------------------------
# longdouble: whole matrices
# ----------------------------------------------------------------------------------------------------------
def _resid_ld(A, C):
    """A - C (C'C)^-1 C'A in longdouble via modified Gram-Schmidt with one re-orthogonalisation (src/wls.jl:221-241
    computes the same projection by QR)."""
    Q = []
    for a in range(C.shape[1]):
        v = C[:, a].copy()
        for _ in range(2):
            for q in Q:
                v = v - q * np.dot(q, v)
        Q.append(v / np.sqrt(np.dot(v, v)))
    R = A.copy()
    for _ in range(2):
        for q in Q:
            R = R - np.outer(q, q @ R)
    return R


def rotate_ld(Ut, *mats):
    Ut = np.asarray(Ut, dtype=LD)
    return [Ut @ np.asarray(M, dtype=LD) for M in mats]


def make_weights_ld(h2, lam):
    """src/lmm.jl:15-33."""
    h2 = LD(h2)
    return LD(1) / (h2 / (LD(1) - h2) * np.asarray(lam, dtype=LD) + LD(1))


def weighted_liteqtl_ld(Y0, G0, C0, lam, h2):
    """src/bulkscan_helpers.jl:175-201 + 47-64 + 22-24 in longdouble: p x m LODs for one shared h2."""
    n = Y0.shape[0]
    sw = np.sqrt(np.abs(make_weights_ld(h2, lam)))
    Y00 = _resid_ld(Y0 * sw[:, None], C0 * sw[:, None])
    X00 = _resid_ld(G0 * sw[:, None], C0 * sw[:, None])
    Y00 = Y00 / np.sqrt(np.sum(Y00 * Y00, axis=0))[None, :]
    X00 = X00 / np.sqrt(np.sum(X00 * X00, axis=0))[None, :]
    R = X00.T @ Y00
    return -(LD(n) / LD(2)) * np.log10(LD(1) - R * R)


def grid_loglik_ld(Y0, C0, lam, grid, prior, reml):
    """wls_multivar(...).Ell per grid point, src/wls.jl:103-176, in longdouble (|grid| x m)."""
    n, c = C0.shape
    a, b = LD(prior[0]), LD(prior[1])
    pdf = b + LD(2) if b > 0 else b
    out = []
    for h in grid:
        w = make_weights_ld(h, lam)
        sw = np.sqrt(w)
        XX = C0 * sw[:, None]
        R = _resid_ld(Y0 * sw[:, None], XX)
        rss0 = np.sum(R * R, axis=0)
        denom = (LD(n - c) if reml else LD(n)) + pdf
        s2 = (rss0 + a * b) / denom
        ll = -LD(0.5) * ((LD(n) + b) * np.log(s2) - np.sum(np.log(w)) + (rss0 + a * b) / s2)
        if reml:
            sign, logdet = np.linalg.slogdet(np.asarray(XX.T @ XX, dtype=np.float64))
            # log det of a c x c SPD matrix: float64 slogdet of the longdouble Gram is accurate to ~1e-15 relative
            ll = ll + LD(0.5) * (LD(c) * np.log(s2) - LD(logdet))
        out.append(ll)
    return np.vstack(out)


def bulkscan_alt_grid_ld(Y, G, C, Ut, lam, grid, prior=(1.0, 0.0), reml=False):
    """bulkscan_alt_grid, src/bulkscan.jl:445-526 with tmax! (src/bulkscan_helpers.jl:330-350), in longdouble.
    Returns (L, h2_panel, logL1 per grid point [K, p, m])."""
    Y0, G0, C0 = rotate_ld(Ut, Y, G, C)
    ln10 = np.log(LD(10))
    ell0 = grid_loglik_ld(Y0, C0, lam, grid, prior, reml)
    prof = []
    for k, h in enumerate(grid):
        prof.append(weighted_liteqtl_ld(Y0, G0, C0, lam, h) * ln10 + ell0[k][None, :])
    prof = np.stack(prof)
    mx = prof[0].copy()
    counter = np.zeros(mx.shape, dtype=np.int64)
    for k in range(1, len(grid)):
        better = mx < prof[k]
        mx = np.where(better, prof[k], mx)
        counter += better
    L = (mx - ell0.max(axis=0)[None, :]) / ln10
    return L, np.asarray(grid)[counter], prof


# ----------------------------------------------------------------------------------------------------------
# mpmath: single entries, effectively exact
# ----------------------------------------------------------------------------------------------------------
def _mp():
    import mpmath
    mpmath.mp.dps = 50
    return mpmath


def _rot_mp(mp, Ut, v):
    n = len(v)
    return [mp.fsum(mp.mpf(float(Ut[a, b])) * mp.mpf(float(v[b])) for b in range(n)) for a in range(n)]


def _project_out_mp(mp, v, cols):
    """v minus its projection on span(cols) (orthonormalised on the fly)."""
    Q = []
    for c in cols:
        u = list(c)
        for q in Q:
            d = mp.fsum(a * b for a, b in zip(q, u))
            u = [a - d * b for a, b in zip(u, q)]
        nrm = mp.sqrt(mp.fsum(a * a for a in u))
        Q.append([a / nrm for a in u])
    r = list(v)
    for q in Q:
        d = mp.fsum(a * b for a, b in zip(q, r))
        r = [a - d * b for a, b in zip(r, q)]
    return r, Q


class TraitMP:
    """One trait and the covariates rotated once in 50-digit arithmetic; markers are rotated on demand."""

    def __init__(self, y, C, Ut, lam):
        self.mp = mp = _mp()
        self.Ut = np.asarray(Ut, dtype=np.float64)
        self.n = len(y)
        self.y0 = _rot_mp(mp, self.Ut, np.asarray(y, dtype=np.float64))
        C = np.asarray(C, dtype=np.float64).reshape(self.n, -1)
        self.C0 = [_rot_mp(mp, self.Ut, C[:, a]) for a in range(C.shape[1])]
        self.lam = [mp.mpf(float(x)) for x in lam]
        self._g0 = {}

    def weights(self, h2):
        mp = self.mp
        h = mp.mpf(float(h2)) if not isinstance(h2, mp.mpf) else h2
        d = h / (1 - h)
        return [1 / (d * l + 1) for l in self.lam]

    def ell(self, h2, prior=(0.0, 0.0), reml=False):
        """wls(...).ell, src/wls.jl:27-97."""
        mp = self.mp
        w = self.weights(h2)
        sw = [mp.sqrt(x) for x in w]
        cols = [[s * v for s, v in zip(sw, c)] for c in self.C0]
        r, Q = _project_out_mp(mp, [s * v for s, v in zip(sw, self.y0)], cols)
        rss0 = mp.fsum(a * a for a in r)
        a, b = mp.mpf(float(prior[0])), mp.mpf(float(prior[1]))
        pdf = b + 2 if b > 0 else b
        n, c = self.n, len(self.C0)
        s2 = (rss0 + a * b) / ((n - c if reml else n) + pdf)
        ll = -(mp.mpf(1) / 2) * ((n + b) * mp.log(s2) - mp.fsum(mp.log(x) for x in w) + (rss0 + a * b) / s2)
        if reml:
            # log|det R|^2 = log det(XX'XX): product of the squared norms Gram-Schmidt removed
            Gm = mp.matrix(c, c)
            for i in range(c):
                for j in range(c):
                    Gm[i, j] = mp.fsum(x * y for x, y in zip(cols[i], cols[j]))
            ll = ll + (mp.mpf(1) / 2) * (c * mp.log(s2) - mp.log(mp.det(Gm)))
        return ll, s2

    def fit_h2(self, prior=(0.0, 0.0), reml=False, lo=0.0, hi=1.0, x0=None):
        """The exact minimiser of -ell on [lo, hi] near x0 (golden section to 1e-20 in 50-digit arithmetic):
        what fitlmm's Brent (src/lmm.jl:56-86) approximates to ~1.5e-8 relative."""
        mp = self.mp
        f = lambda h: -self.ell(h, prior, reml)[0]
        a, b = mp.mpf(lo), mp.mpf(hi)
        if x0 is not None:  # bracket around the float64 answer
            a, b = max(a, mp.mpf(x0) - mp.mpf("1e-4")), min(b, mp.mpf(x0) + mp.mpf("1e-4"))
        b = min(b, 1 - mp.mpf("1e-30"))
        gr = (mp.sqrt(5) - 1) / 2
        c, d = b - gr * (b - a), a + gr * (b - a)
        fc, fd = f(c), f(d)
        while b - a > mp.mpf("1e-18"):
            if fc < fd:
                b, d, fd = d, c, fc
                c = b - gr * (b - a)
                fc = f(c)
            else:
                a, c, fc = c, d, fd
                d = a + gr * (b - a)
                fd = f(d)
        return (a + b) / 2

    def lod(self, g, h2, use_abs=True):
        """weighted_liteqtl for one marker, src/bulkscan_helpers.jl:175-201 (use_abs: sqrt(abs(w)) as there)."""
        mp = self.mp
        g0 = _rot_mp(mp, self.Ut, np.asarray(g, dtype=np.float64))
        w = self.weights(h2)
        sw = [mp.sqrt(abs(x)) if use_abs else mp.sqrt(x) for x in w]
        cols = [[s * v for s, v in zip(sw, c)] for c in self.C0]
        ry, _ = _project_out_mp(mp, [s * v for s, v in zip(sw, self.y0)], cols)
        rg, _ = _project_out_mp(mp, [s * v for s, v in zip(sw, g0)], cols)
        num = mp.fsum(a * b for a, b in zip(ry, rg))
        r2 = num * num / (mp.fsum(a * a for a in ry) * mp.fsum(a * a for a in rg))
        return -(mp.mpf(self.n) / 2) * mp.log10(1 - r2)

    def alt_grid_entry(self, g, grid: Sequence[float], prior=(1.0, 0.0), reml=False):
        """(L, h2_panel value, profile) of one (marker, trait) entry of bulkscan_alt_grid."""
        mp = self.mp
        ln10 = mp.log(10)
        ell0 = [self.ell(h, prior, reml)[0] for h in grid]
        prof = [self.lod(g, h) * ln10 + e for h, e in zip(grid, ell0)]
        mx, cnt = prof[0], 0
        for k in range(1, len(grid)):
            if mx < prof[k]:
                mx = prof[k]
                cnt += 1
        return (mx - max(ell0)) / ln10, grid[cnt], prof


def closer_side(truth, a, b):
    """(-1 if a is closer to `truth` (an mpmath value), +1 if b is, 0 on a tie, |truth - a|, |truth - b|)."""
    mp = _mp()
    ea = float(abs(truth - mp.mpf(float(a))))
    eb = float(abs(truth - mp.mpf(float(b))))
    return (-1 if ea < eb else (1 if eb < ea else 0)), ea, eb
