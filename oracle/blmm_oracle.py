"""CPU oracle: a numpy float64 restatement of BulkLMM.jl's multi-trait LMM genome-scan path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`bulklmm.jl_b200/`) imports this
module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may.  It is the checker, never the thing shipped.

Pinning status (see DESIGN.md "Oracle"):
  * pinned by the reference's own arithmetic known-answer tests that need no data files
    (`test/bulkscan_test.jl:9-19` r2lod inverse, `test/gridbrent_test.jl:2-8`,
    `test/lmm_test.jl:12-18` error string) and by every data-independent identity the
    reference's tests assert between its own code paths (tests/test_oracle_*.py);
  * PARITY UNPINNED against reference *outputs*: Julia is not installed in the build
    container, and the BXD genotype/phenotype CSVs that every data-driven reference test
    reads are absent from the checkout (`/root/reference/.MISSING_LARGE_BLOBS`), so the
    lmmlite/GEMMA golden LODs shipped with the reference cannot be replayed.
  * third-party arithmetic restated from its published algorithm: Optim.jl `Brent()`
    (version unpinned by the reference: `Project.toml` compat "1.7, 2"), LAPACK `dsyevr`
    via `scipy.linalg.eigh(driver="evr")` (what Julia's `eigen(::Symmetric-valued Matrix)` calls).

Each function cites the reference file:line it follows (paths relative to /root/reference).
All arrays are float64; matrices are n x m (traits), n x p (markers), p x m (LOD) exactly as in
the reference (Julia column-major is irrelevant to numpy semantics).
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional, Sequence

import numpy as np
import scipy.linalg as sla

LN10 = math.log(10.0)


class OracleError(Exception):
    """Mirrors Julia `error(msg)`; `.msg` holds the reference's exact string."""

    def __init__(self, msg: str):
        super().__init__(msg)
        self.msg = msg


# ----------------------------------------------------------------------------------------
# util.jl
# ----------------------------------------------------------------------------------------
def check_zeros(x: np.ndarray) -> bool:
    """src/util.jl:47-56 — isapprox(x_i, 0; atol=eps, rtol=0) for any i."""
    return bool(np.any(np.abs(x) <= np.finfo(np.float64).eps))


def col_divide(A: np.ndarray, x: np.ndarray) -> np.ndarray:
    """src/util.jl:58-78 colDivide! (returns the divided matrix instead of mutating)."""
    if x.shape[0] != A.shape[1]:
        raise OracleError("Matrix and vector size do not match.")
    if check_zeros(x):
        raise OracleError("Dividing by zeros: the input vector can not contain any zeros!")
    return A / x[None, :]


def row_multiply(A: np.ndarray, x: np.ndarray) -> np.ndarray:
    """src/util.jl:139-156 rowMultiply."""
    if x.shape[0] != A.shape[0]:
        raise OracleError("Matrix and vector size do not match.")
    return A * x[:, None]


def col_standardize(A: np.ndarray) -> np.ndarray:
    """src/util.jl:88-96 colStandardize (sample std, ddof=1 as Julia `std`)."""
    sA = A - A.mean(axis=0, keepdims=True)
    return col_divide(sA, sA.std(axis=0, ddof=1))


def shuffle_vector(x: np.ndarray, perm_idx: np.ndarray, original: bool = True) -> np.ndarray:
    """src/util.jl:162-179 shuffleVector with the shuffles given as index columns.

    The reference draws `shuffle(rng, x)` from `MersenneTwister(rndseed)`; that stream is
    Julia-version dependent and cannot be reproduced here, so the permutation *indices*
    (n x nshuffle, 0-based: column s of the result is x[perm_idx[:, s]]) are an input.
    """
    n, ns = perm_idx.shape
    cols = [x] if original else []
    out = np.empty((n, ns + (1 if original else 0)))
    if original:
        out[:, 0] = x
    out[:, (1 if original else 0):] = x[perm_idx]
    return out


def make_perm_indices(n: int, nperms: int, rndseed: int = 0) -> np.ndarray:
    """Stand-in for MersenneTwister(rndseed)+shuffle (src/transform_helpers.jl:98, src/util.jl:175)."""
    rng = np.random.default_rng(rndseed)
    idx = np.tile(np.arange(n, dtype=np.int32)[:, None], (1, nperms))
    return rng.permuted(idx, axis=0)


def _logccdf_chisq_tail(x: np.ndarray, df: int) -> np.ndarray:
    """log Q(df/2, x/2) for large x, where chi2.logsf = log(sf) underflows to -inf (x > ~1480).
    Q(a+1, y) = Q(a, y) + y^a e^-y / Gamma(a+1) in log space, from Q(1/2, y) = erfc(sqrt y) = 2 Phi(-sqrt(2y))
    (log_ndtr keeps full accuracy in the tail) or Q(1, y) = e^-y."""
    from scipy.special import gammaln, log_ndtr

    y = 0.5 * np.asarray(x, dtype=np.float64)
    if df % 2:
        a, lq = 0.5, np.log(2.0) + log_ndtr(-np.sqrt(2.0 * y))
    else:
        a, lq = 1.0, -y
    while a < 0.5 * df - 0.25:
        lq = np.logaddexp(lq, a * np.log(y) - y - gammaln(a + 1.0))
        a += 1.0
    return lq


def lod2log10p(lod, df: int = 1):
    """src/util.jl:199-206: -logccdf(Chisq(df), 2 ln10 lod)/ln10.  Distributions.jl evaluates logccdf in log space;
    scipy's logsf is log(sf) and underflows beyond LOD ~ 320, so the far tail is restated explicitly."""
    from scipy.stats import chi2

    x = np.asarray(lod, dtype=np.float64) * 2.0 * LN10
    ls = np.asarray(chi2.logsf(x, df), dtype=np.float64)
    far = np.isfinite(x) & (x > 1000.0)
    if np.any(far):
        ls = ls.copy()
        ls[far] = _logccdf_chisq_tail(x[far], df)
    return -ls / LN10


def lod2p(lod, df: int = 1):
    """src/util.jl:190-197."""
    from scipy.stats import chi2

    return chi2.sf(np.asarray(lod) * 2.0 * LN10, df)


def p2lod(pval, df: int = 1):
    """src/util.jl:181-188."""
    from scipy.stats import chi2

    return chi2.isf(np.asarray(pval), df) / (2.0 * LN10)


# ----------------------------------------------------------------------------------------
# kinship.jl
# ----------------------------------------------------------------------------------------
def calc_kinship(geno: np.ndarray) -> np.ndarray:
    """src/kinship.jl:4-14."""
    X = geno - 0.5
    K = 2.0 * (X @ X.T) / X.shape[1] + 0.5
    np.fill_diagonal(K, 1.0)
    return K


# ----------------------------------------------------------------------------------------
# wls.jl
# ----------------------------------------------------------------------------------------
class LSEstimates(NamedTuple):
    b: np.ndarray
    sigma2: float
    ell: float


class LSEstimatesMultivar(NamedTuple):
    B: np.ndarray
    Sigma2: np.ndarray
    Ell: np.ndarray


def _ls_solve(XX: np.ndarray, YY: np.ndarray, method: str):
    """The two factorisations of src/wls.jl:49-67 / 126-143; returns (coef, logdet(XX'XX))."""
    if method == "cholesky":
        S = XX.T @ XX
        L = np.linalg.cholesky(S)
        coef = sla.cho_solve((L, True), XX.T @ YY)
        logdet = 2.0 * np.sum(np.log(np.diag(L)))
    elif method == "qr":
        Q, R = np.linalg.qr(XX, mode="reduced")
        coef = sla.solve_triangular(R, Q.T @ YY, lower=False)
        logdet = 2.0 * np.sum(np.log(np.abs(np.diag(R))))
    else:
        raise OracleError("unknown method")
    return coef, logdet


def _prior_df(prior) -> float:
    """src/wls.jl:72-76."""
    return prior[1] + 2.0 if prior[1] > 0.0 else prior[1]


def wls(y, X, w, prior, reml=False, loglik=True, method="qr") -> LSEstimates:
    """src/wls.jl:27-97."""
    out = wls_multivar(y, X, w, prior, reml=reml, loglik=loglik, method=method)
    ell = float(out.Ell[0, 0]) if loglik else float("nan")
    return LSEstimates(out.B, float(out.Sigma2[0, 0]), ell)


def wls_multivar(Y, X, w, prior, reml=False, loglik=True, method="qr") -> LSEstimatesMultivar:
    """src/wls.jl:103-176."""
    n, p = X.shape
    n = Y.shape[0]
    sqrtw = np.sqrt(w)
    YY = row_multiply(Y, sqrtw)
    XX = row_multiply(X, sqrtw)
    coef, logdet = _ls_solve(XX, YY, method)
    YYhat = XX @ coef
    rss0 = (np.linalg.norm(YY - YYhat, axis=0) ** 2)[None, :]
    pdf = _prior_df(prior)
    ab = prior[0] * prior[1]
    if reml:
        sigma2 = (rss0 + ab) / ((n - p) + pdf)
    else:
        sigma2 = (rss0 + ab) / (n + pdf)
    if loglik:
        ll = -0.5 * ((n + prior[1]) * np.log(sigma2) - np.sum(np.log(w)) + (rss0 + ab) / sigma2)
        if reml:
            ll = ll + 0.5 * (p * np.log(sigma2) - logdet)
    else:
        ll = np.full_like(sigma2, np.nan)
    return LSEstimatesMultivar(coef, sigma2, ll)


def resid(y, X, method="qr"):
    """src/wls.jl:221-263."""
    X = X.reshape(X.shape[0], -1)
    if method == "cholesky":
        b = np.linalg.solve(X.T @ X, X.T @ y)
    else:
        Q, R = np.linalg.qr(X, mode="reduced")
        b = sla.solve_triangular(R, Q.T @ y, lower=False)
    return y - X @ b


def rss(y, X, method="qr"):
    """src/wls.jl:191-207."""
    r = resid(y, X, method=method)
    return np.sum(r**2, axis=0, keepdims=True)


# ----------------------------------------------------------------------------------------
# lmm.jl, gridbrent.jl
# ----------------------------------------------------------------------------------------
def make_weights(h2: float, lam: np.ndarray) -> np.ndarray:
    """src/lmm.jl:15-33."""
    denom = 1.0 - h2
    delta = math.inf if denom == 0.0 else h2 / denom
    if math.isinf(delta):
        raise OracleError("Heritability of 1 is not allowed.")
    return 1.0 / (delta * lam + 1.0)


class BrentResult(NamedTuple):
    minimizer: float
    minimum: float
    f_calls: int


def brent_minimize(f, lo: float, hi: float, rel_tol: float = math.sqrt(np.finfo(float).eps),
                   abs_tol: float = np.finfo(float).eps, iterations: int = 1000) -> BrentResult:
    """Optim.jl `optimize(f, lo, hi, Brent())` (call site src/gridbrent.jl:16).

    Optim.jl is not vendored in /root/reference; this restates its published univariate
    Brent solver (golden-section start point, parabolic step with the 2*x_tol guards, the
    three-point bookkeeping) with Optim's default tolerances rel_tol=sqrt(eps), abs_tol=eps.
    """
    golden = 0.5 * (3.0 - math.sqrt(5.0))
    x = lo + golden * (hi - lo)
    fx = f(x)
    calls = 1
    step = 0.0
    old_step = 0.0
    xo = xoo = x
    fo = foo = fx
    it = 0
    while it < iterations:
        p = 0.0
        q = 0.0
        tol = rel_tol * abs(x) + abs_tol
        mid = (hi + lo) / 2.0
        if abs(x - mid) <= 2.0 * tol - (hi - lo) / 2.0:
            break
        it += 1
        if abs(old_step) > tol:
            r = (x - xo) * (fx - foo)
            q = (x - xoo) * (fx - fo)
            p = (x - xoo) * q - (x - xo) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            else:
                q = -q
        if abs(p) < abs(q * old_step / 2.0) and p < q * (hi - x) and p < q * (x - lo):
            old_step = step
            step = p / q
            xt = x + step
            if (xt - lo) < 2.0 * tol or (hi - xt) < 2.0 * tol:
                step = tol if x < mid else -tol
        else:
            old_step = (hi - x) if x < mid else (lo - x)
            step = golden * old_step
        if abs(step) >= tol:
            xn = x + step
        else:
            xn = x + (tol if step > 0.0 else -tol)
        fn = f(xn)
        calls += 1
        if fn < fx:
            if xn < x:
                hi = x
            else:
                lo = x
            xoo, foo = xo, fo
            xo, fo = x, fx
            x, fx = xn, fn
        else:
            if xn < x:
                lo = xn
            else:
                hi = xn
            if fn <= fo or xo == x:
                xoo, foo = xo, fo
                xo, fo = xn, fn
            elif fn <= foo or xoo == x or xoo == xo:
                xoo, foo = xn, fn
    return BrentResult(x, fx, calls)


def gridbrent(f, a: float, b: float, ninterval: int = 1) -> BrentResult:
    """src/gridbrent.jl:9-24 (argmin keeps the first of equal minima)."""
    pts = np.linspace(a, b, ninterval + 1)
    res = [brent_minimize(f, float(pts[i]), float(pts[i + 1])) for i in range(ninterval)]
    idx = int(np.argmin([r.minimum for r in res]))
    return res[idx]


class LMMEstimates(NamedTuple):
    b: np.ndarray
    sigma2: float
    h2: float
    ell: float


def fitlmm(y, X, lam, prior, reml=False, loglik=True, method="qr", optim_interval=1,
           h20=0.5, d=1.0) -> LMMEstimates:
    """src/lmm.jl:56-86."""

    def neg_ll(h2):
        return -wls(y, X, make_weights(h2, lam), prior, reml=reml, loglik=loglik, method=method).ell

    lb = max(h20 - d, 0.0)
    ub = min(h20 + d, 1.0)
    opt = gridbrent(neg_ll, lb, ub, optim_interval)
    h2 = opt.minimizer
    est = wls(y, X, make_weights(h2, lam), prior, reml=reml, loglik=loglik, method=method)
    return LMMEstimates(est.b, est.sigma2, h2, est.ell)


# ----------------------------------------------------------------------------------------
# transform_helpers.jl
# ----------------------------------------------------------------------------------------
def decompose(K: np.ndarray, decomp_scheme: str = "eigen"):
    """The factorisation inside src/transform_helpers.jl:21-49: returns (Ut, values)."""
    if decomp_scheme == "eigen":
        vals, vecs = sla.eigh(K, driver="evr")
        return vecs.T.copy(), vals
    if decomp_scheme == "svd":
        _, S, Vt = np.linalg.svd(K)
        return Vt, S
    raise OracleError("Please choose either `eigen` or `svd` for decomposition of the kinship matrix.")


def transform_rotation(y, g, K, addIntercept=True, decomp_scheme="eigen", Ut=None, lam=None):
    """src/transform_helpers.jl:1-54.  A precomputed (Ut, lam) may be supplied so the oracle
    and the engine rotate with the *same* eigenvectors (permutation LODs depend on their signs)."""
    n = y.shape[0]
    if g.shape[0] != n or K.shape[0] != n:
        raise OracleError("Dimension mismatch.")
    X = np.hstack([np.ones((n, 1)), g]) if addIntercept else g
    if Ut is None:
        Ut, lam = decompose(K, decomp_scheme)
    return Ut @ y, Ut @ X, np.asarray(lam, dtype=np.float64)


def transform_reweight(y0, X0, lam, n_covars=1, prior_a=0.0, prior_b=0.0, method="qr",
                       optim_interval=1, reml=False):
    """src/transform_helpers.jl:57-92."""
    vc = fitlmm(y0, X0[:, :n_covars], lam, [prior_a, prior_b], reml=reml, method=method,
                optim_interval=optim_interval)
    r0 = y0 - X0[:, :n_covars] @ vc.b
    sqrtw = np.sqrt(make_weights(vc.h2, lam))
    r0w = row_multiply(r0, sqrtw)
    X0w = row_multiply(X0, sqrtw)
    X00 = resid(X0w[:, n_covars:], X0w[:, :n_covars])
    return r0w, X00, vc.sigma2, vc.h2


def transform_permute(r0, perm_idx: np.ndarray, original=True):
    """src/transform_helpers.jl:94-102 with explicit permutation indices (see shuffle_vector)."""
    return shuffle_vector(r0[:, 0], perm_idx, original=original)


# ----------------------------------------------------------------------------------------
# bulkscan_helpers.jl
# ----------------------------------------------------------------------------------------
def r2lod(r, n: int):
    """src/bulkscan_helpers.jl:22-24."""
    return -(n / 2.0) * np.log10(1.0 - np.asarray(r) ** 2)


def compute_r_lmm(wY, wX, wIntercept):
    """src/bulkscan_helpers.jl:47-64."""
    Y00 = resid(wY, wIntercept)
    X00 = resid(wX, wIntercept)
    norm_Y = np.linalg.norm(Y00, axis=0)
    norm_X = np.linalg.norm(X00, axis=0)
    Y00 = col_divide(Y00, norm_Y)
    X00 = col_divide(X00, norm_X)
    return X00.T @ Y00


def weighted_liteqtl(Y0, X0, lam, hsq, num_of_covar=1):
    """src/bulkscan_helpers.jl:175-201."""
    n = Y0.shape[0]
    sqrtw = np.sqrt(np.abs(make_weights(hsq, lam)))
    wY0 = row_multiply(Y0, sqrtw)
    wX0 = row_multiply(X0, sqrtw)
    R = compute_r_lmm(wY0, wX0[:, num_of_covar:], wX0[:, :num_of_covar])
    return r2lod(R, n)


def univar_liteqtl(y0_j, X0_intercept, X0_covar, lam, prior_variance=0.0, prior_sample_size=0.0,
                   reml=False, optim_interval=1):
    """src/bulkscan_helpers.jl:127-150 (X0_intercept = covariate block, X0_covar = marker block)."""
    n = y0_j.shape[0]
    y0 = y0_j.reshape(-1, 1)
    vc = fitlmm(y0, X0_intercept, lam, [prior_variance, prior_sample_size], reml=reml,
                optim_interval=optim_interval)
    sqrtw = np.sqrt(np.abs(make_weights(vc.h2, lam)))
    R = compute_r_lmm(row_multiply(y0, sqrtw), row_multiply(X0_covar, sqrtw),
                      row_multiply(X0_intercept, sqrtw))
    return r2lod(R, n), vc.h2


def find_optim_h2(h2_list, results):
    """src/bulkscan_helpers.jl:204-211 (`findmax` => first maximum on ties)."""
    return np.asarray(h2_list)[np.argmax(results, axis=0)]


def grid_loglik(Y0, X0_cov, lam, grid, prior, reml=False):
    """The ell_results matrix of src/bulkscan_helpers.jl:267-269 (|grid| x m)."""
    return np.vstack([wls_multivar(Y0, X0_cov, make_weights(h, lam), prior, reml=reml).Ell for h in grid])


def gridscan_by_bin(pheno, geno, covar, kinship, grid, addIntercept=True, prior_variance=1.0,
                    prior_sample_size=0.0, reml=False, decomp_scheme="eigen", Ut=None, lam=None):
    """src/bulkscan_helpers.jl:239-292.  Returns (masks per bin, LODs per bin, h2 per bin)."""
    m = pheno.shape[1]
    Y0, X0, lam0 = transform_rotation(pheno, np.hstack([covar, geno]), kinship,
                                      addIntercept=addIntercept, decomp_scheme=decomp_scheme,
                                      Ut=Ut, lam=lam)
    prior = [prior_variance, prior_sample_size]
    c = covar.shape[1] + (1 if addIntercept else 0)
    ell = grid_loglik(Y0, X0[:, :c], lam0, grid, prior, reml=reml)
    optim_h2 = find_optim_h2(grid, ell)
    # `unique(values(Dict))` order is hash order in Julia; the bin order does not affect the
    # reassembled result, so first-appearance order is used here.
    h2_taken = list(dict.fromkeys(optim_h2.tolist()))
    masks = [optim_h2 == h for h in h2_taken]
    lods = [weighted_liteqtl(Y0[:, mk], X0, lam0, h, num_of_covar=c) for mk, h in zip(masks, h2_taken)]
    return masks, lods, h2_taken


class NullScan(NamedTuple):
    L: np.ndarray
    h2_null_list: np.ndarray


class AltScan(NamedTuple):
    L: np.ndarray
    h2_panel: np.ndarray


def _apply_obs_weights(Y, G, Covar, K, weights, addIntercept):
    """The `weights` pre-scaling block shared by src/bulkscan.jl:231-250, 351-370, 457-476."""
    if weights is None:
        return Y, G, Covar, K, addIntercept
    W = np.asarray(weights, dtype=np.float64)
    Y_st = W[:, None] * Y
    G_st = W[:, None] * G
    if addIntercept:
        Covar_st = W[:, None] * np.hstack([np.ones((Y.shape[0], 1)), Covar])
    else:
        Covar_st = W[:, None] * Covar
    K_st = W[:, None] * K * W[None, :]
    return Y_st, G_st, Covar_st, K_st, False


def _default_covar(Y, Covar, addIntercept):
    """3-argument forms: intercept is the only covariate, addIntercept=false
    (src/bulkscan.jl:94-109, 200-209, 327-337, 434-442)."""
    if Covar is None:
        return np.ones((Y.shape[0], 1)), False
    return Covar, addIntercept


def bulkscan_null_grid(Y, G, K, grid_list, Covar=None, weights=None, addIntercept=True,
                       prior_variance=1.0, prior_sample_size=0.0, reml=False,
                       decomp_scheme="eigen", Ut=None, lam=None) -> NullScan:
    """src/bulkscan.jl:321-385 (+ reorder_results src/bulkscan_helpers.jl:294-308,
    get_h2_distribution src/bulkscan.jl:387-397)."""
    Covar, addIntercept = _default_covar(Y, Covar, addIntercept)
    m, p = Y.shape[1], G.shape[1]
    Y_st, G_st, Covar_st, K_st, addIntercept = _apply_obs_weights(Y, G, Covar, K, weights, addIntercept)
    masks, lods, h2_taken = gridscan_by_bin(Y_st, G_st, Covar_st, K_st, np.asarray(grid_list, float),
                                            addIntercept=addIntercept, prior_variance=prior_variance,
                                            prior_sample_size=prior_sample_size, reml=reml,
                                            decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    L = np.empty((p, m))
    h2 = np.zeros(m)
    for mk, lod, h in zip(masks, lods, h2_taken):
        L[:, mk] = lod
        h2[mk] = h
    return NullScan(L, h2)


def tmax(mx, to_compare, hsq_panel, counter, hsq_list):
    """src/bulkscan_helpers.jl:330-350 tmax! — NOTE the counter semantics: on every
    strict improvement the counter advances by ONE and h2_panel := hsq_list[counter]; it is not
    the arg-max index."""
    upd = mx < to_compare
    mx[upd] = to_compare[upd]
    counter[upd] += 1
    hsq_panel[upd] = np.asarray(hsq_list)[counter[upd] - 1]  # counter is 1-based as in Julia


def bulkscan_alt_grid(Y, G, K, hsq_list, Covar=None, reml=False, prior_variance=1.0,
                      prior_sample_size=0.0, weights=None, addIntercept=True,
                      decomp_scheme="eigen", Ut=None, lam=None, profile: Optional[list] = None) -> AltScan:
    """src/bulkscan.jl:428-526.

    `profile` (not in the reference): a list that receives the p x m matrix `logL1_k` of every grid point in order —
    the operands of tmax!'s strict comparison; tests use them to show that an h2_panel disagreement is a
    rounding-level tie of two of them and nothing else.

    Divergence from the reference, on purpose: the grid loop at src/bulkscan.jl:510 omits
    `num_of_covar`, which makes the reference throw DimensionMismatch whenever c > 1 (SURVEY
    B1).  The intended arithmetic (pass c at every grid point) is restated here; with c == 1
    — the only case the reference can run — the two are identical.
    """
    Covar, addIntercept = _default_covar(Y, Covar, addIntercept)
    p, m = G.shape[1], Y.shape[1]
    n_cov_in = Covar.shape[1]
    add0 = addIntercept
    Y_st, G_st, Covar_st, K_st, addIntercept = _apply_obs_weights(Y, G, Covar, K, weights, addIntercept)
    Y0, X0, lam0 = transform_rotation(Y_st, np.hstack([Covar_st, G_st]), K_st, addIntercept=addIntercept,
                                      decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    # intended covariate count (the reference's own count is off by one with weights+addIntercept)
    c = n_cov_in + (1 if add0 else 0)
    X0_base = X0[:, :c]
    prior = [prior_variance, prior_sample_size]
    hsq_list = [float(h) for h in hsq_list]

    logLR = weighted_liteqtl(Y0, X0, lam0, hsq_list[0], num_of_covar=c) * LN10
    logL0 = wls_multivar(Y0, X0_base, make_weights(hsq_list[0], lam0), prior, reml=reml).Ell
    logL1 = logLR + np.repeat(logL0, p, axis=0)
    if profile is not None:
        profile.append(logL1.copy())
    logL0_all = np.zeros((len(hsq_list), m))
    logL0_all[0, :] = logL0
    h2_panel = np.ones((p, m)) * hsq_list[0]
    counter = np.ones((p, m), dtype=np.int64)
    for k, h2 in enumerate(hsq_list[1:], start=1):
        logLR_k = weighted_liteqtl(Y0, X0, lam0, h2, num_of_covar=c) * LN10
        logL0_k = wls_multivar(Y0, X0_base, make_weights(h2, lam0), prior, reml=reml).Ell
        logL1_k = logLR_k + np.repeat(logL0_k, p, axis=0)
        logL0_all[k, :] = logL0_k
        if profile is not None:
            profile.append(logL1_k)
        tmax(logL1, logL1_k, h2_panel, counter, hsq_list)
    logL0_opt = np.max(logL0_all, axis=0, keepdims=True)
    L = (logL1 - np.repeat(logL0_opt, p, axis=0)) / LN10
    return AltScan(L, h2_panel)


def bulkscan_null(Y, G, K, Covar=None, nb=1, weights=None, addIntercept=True, prior_variance=1.0,
                  prior_sample_size=0.0, reml=False, optim_interval=1, decomp_scheme="eigen",
                  Ut=None, lam=None, h2_override: Optional[np.ndarray] = None) -> NullScan:
    """src/bulkscan.jl:188-314 (`nb`/`nt_blas` only shape the threading, not the result).

    `h2_override` (not in the reference) scans with given per-trait h2 instead of the Brent fit;
    tests use it to check LODs at the engine's own h2 estimates (two independent FP64 Brent
    implementations cannot agree to 1e-8, SURVEY section 7 hard part 3a).
    """
    Covar, addIntercept = _default_covar(Y, Covar, addIntercept)
    m, p = Y.shape[1], G.shape[1]
    c = Covar.shape[1] + (1 if addIntercept else 0)
    Y_st, G_st, Covar_st, K_st, addIntercept = _apply_obs_weights(Y, G, Covar, K, weights, addIntercept)
    Y0, X0, lam0 = transform_rotation(Y_st, np.hstack([Covar_st, G_st]), K_st, addIntercept=addIntercept,
                                      decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    X0_cov = X0[:, :c]
    X0_mark = X0[:, c:]
    L = np.empty((p, m))
    h2s = np.zeros(m)
    for j in range(m):
        if h2_override is None:
            lod, h2 = univar_liteqtl(Y0[:, j], X0_cov, X0_mark, lam0, prior_variance=prior_variance,
                                     prior_sample_size=prior_sample_size, reml=reml,
                                     optim_interval=optim_interval)
        else:
            h2 = float(h2_override[j])
            lod = weighted_liteqtl(Y0[:, j:j + 1], X0, lam0, h2, num_of_covar=c)
        L[:, j] = lod[:, 0]
        h2s[j] = h2
    return NullScan(L, h2s)


def bulkscan(Y, G, K, Covar=None, method="null-grid", h2_grid=None, nb=1, nt_blas=1, addIntercept=True,
             weights=None, prior_variance=1.0, prior_sample_size=0.0, reml=False, optim_interval=1,
             decomp_scheme="eigen", output_pvals=False, chisq_df=1, Ut=None, lam=None):
    """src/bulkscan.jl:81-162 dispatcher.  Returns a dict with the NamedTuple's field names."""
    if h2_grid is None:
        h2_grid = np.arange(10) / 10.0  # collect(0.0:0.1:0.9)
    kw = dict(Covar=Covar, weights=weights, addIntercept=addIntercept, prior_variance=prior_variance,
              prior_sample_size=prior_sample_size, reml=reml, decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    if method == "null-exact":
        r = bulkscan_null(Y, G, K, nb=nb, optim_interval=optim_interval, **kw)
        out = {"L": r.L, "h2_null_list": r.h2_null_list}
    elif method == "null-grid":
        r = bulkscan_null_grid(Y, G, K, h2_grid, **kw)
        out = {"L": r.L, "h2_null_list": r.h2_null_list}
    elif method == "alt-grid":
        r = bulkscan_alt_grid(Y, G, K, h2_grid, **kw)
        out = {"L": r.L, "h2_panel": r.h2_panel}
    else:
        raise OracleError("unknown method")
    if output_pvals:
        out["log10Pvals_mat"] = lod2log10p(out["L"], chisq_df)
        out["Chisq_df"] = chisq_df
    return out


# ----------------------------------------------------------------------------------------
# scan.jl
# ----------------------------------------------------------------------------------------
def scan_null(y, g, covar, K, prior, addIntercept, reml=False, method="qr", optim_interval=1,
              decomp_scheme="eigen", Ut=None, lam=None):
    """src/scan.jl:310-360 — per-marker QR rss loop (no abs() on the weights here)."""
    n, p = g.shape
    c = covar.shape[1] + (1 if addIntercept else 0)
    y0, X0, lam0 = transform_rotation(y, np.hstack([covar, g]), K, addIntercept=addIntercept,
                                      decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    X0_cov = X0[:, :c]
    out00 = fitlmm(y0, X0_cov, lam0, prior, reml=reml, method=method, optim_interval=optim_interval)
    sqrtw = np.sqrt(make_weights(out00.h2, lam0))
    y0w = row_multiply(y0, sqrtw)
    X0w = row_multiply(X0, sqrtw)
    rss0 = rss(y0w, X0w[:, :c], method=method)[0, 0]
    lod = np.zeros(p)
    X = X0w[:, :c + 1].copy()
    for i in range(p):
        X[:, c] = X0w[:, c + i]
        rss1 = rss(y0w, X, method=method)[0, 0]
        lod[i] = (-n / 2.0) * (math.log10(rss1) - math.log10(rss0))
    return {"sigma2_e": out00.sigma2, "h2_null": out00.h2, "lod": lod}


def scan_alt(y, g, covar, K, prior, addIntercept, reml=False, method="qr", optim_interval=1,
             decomp_scheme="eigen", Ut=None, lam=None):
    """src/scan.jl:397-453 — variance components re-estimated for every marker.  As in the reference the two
    final likelihoods are `wls(..., sqrtw, prior)`: the SQUARE ROOTS of the weights are handed to wls as its
    weights, and reml is left at wls's default (false) whatever `reml` was used for the fits."""
    n, p = g.shape
    c = covar.shape[1] + (1 if addIntercept else 0)
    y0, X0, lam0 = transform_rotation(y, np.hstack([covar, g]), K, addIntercept=addIntercept,
                                      decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    X0_cov = X0[:, :c]
    out00 = fitlmm(y0, X0_cov, lam0, prior, reml=reml, method=method, optim_interval=optim_interval)
    sqrtw_null = np.sqrt(make_weights(out00.h2, lam0))
    ell_null = wls(y0, X0_cov, sqrtw_null, prior).ell
    lod = np.zeros(p)
    pve = np.zeros(p)
    X = X0[:, :c + 1].copy()
    for i in range(p):
        X[:, c] = X0[:, c + i]
        out11 = fitlmm(y0, X, lam0, prior, reml=reml, method=method, optim_interval=optim_interval)
        sqrtw_alt = np.sqrt(make_weights(out11.h2, lam0))
        lod[i] = (wls(y0, X, sqrtw_alt, prior).ell - ell_null) / math.log(10.0)
        pve[i] = out11.h2
    return {"sigma2_e": out00.sigma2, "h2_null": out00.h2, "h2_each_marker": pve, "lod": lod}


def scan_perms_lite(y, g, covar, K, perm_idx, prior_variance=1.0, prior_sample_size=0.0,
                    addIntercept=True, method="qr", optim_interval=1, reml=False,
                    decomp_scheme="eigen", Ut=None, lam=None):
    """src/scan.jl:485-557 with the shuffles supplied as indices (n x nperms, 0-based)."""
    if y.shape[1] != 1:
        raise OracleError("Can only handle one trait.")
    n = g.shape[0]
    y0, X0, lam0 = transform_rotation(y, np.hstack([covar, g]), K, addIntercept=addIntercept,
                                      decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    c = covar.shape[1] + (1 if addIntercept else 0)
    r0, X00, sigma2_e, h2_null = transform_reweight(y0, X0, lam0, n_covars=c, prior_a=prior_variance,
                                                    prior_b=prior_sample_size, reml=reml, method=method,
                                                    optim_interval=optim_interval)
    r0perm = transform_permute(r0, perm_idx, original=True)
    r0perm = col_divide(r0perm, np.linalg.norm(r0perm, axis=0))
    X00 = col_divide(X00, np.linalg.norm(X00, axis=0))
    L = r2lod(X00.T @ r0perm, n)
    return {"sigma2_e": sigma2_e, "h2_null": h2_null, "lod": L[:, 0].copy(), "L_perms": L[:, 1:].copy()}


def scan(y, g, K, covar=None, weights=None, prior_variance=0.0, prior_sample_size=0.0,
         addIntercept=True, reml=False, assumption="null", method="qr", optim_interval=1,
         permutation_test=False, nperms=1024, rndseed=0, perm_idx=None, decomp_scheme="eigen",
         Ut=None, lam=None):
    """src/scan.jl:94-271."""
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y.reshape(-1, 1)
    n = y.shape[0]
    if covar is None:
        if not addIntercept:
            raise OracleError("Intercept has to be added when no other covariate is given.")
        covar = np.ones((n, 1))
        addIntercept = False
    if weights is not None:
        W = np.asarray(weights, dtype=np.float64)
        y = W[:, None] * y
        g = W[:, None] * g
        covar = W[:, None] * (np.hstack([np.ones((n, 1)), covar]) if addIntercept else covar)
        K = W[:, None] * K * W[None, :]
        addIntercept = False
    if assumption == "alt":
        if permutation_test:
            raise OracleError("Permutation test option currently is not supported for the alternative assumption.")
        return scan_alt(y, g, covar, K, [prior_variance, prior_sample_size], addIntercept, reml=reml,
                        method=method, optim_interval=optim_interval, decomp_scheme=decomp_scheme, Ut=Ut, lam=lam)
    if assumption != "null":
        raise OracleError("Assumption keyword is not supported. Please enter null or alt.")
    if permutation_test:
        if perm_idx is None:
            perm_idx = make_perm_indices(n, nperms, rndseed)
        return scan_perms_lite(y, g, covar, K, perm_idx, prior_variance=prior_variance,
                               prior_sample_size=prior_sample_size, addIntercept=addIntercept, reml=reml,
                               method=method, optim_interval=optim_interval, decomp_scheme=decomp_scheme,
                               Ut=Ut, lam=lam)
    return scan_null(y, g, covar, K, [prior_variance, prior_sample_size], addIntercept, reml=reml,
                     method=method, optim_interval=optim_interval, decomp_scheme=decomp_scheme,
                     Ut=Ut, lam=lam)


# ----------------------------------------------------------------------------------------
# analysis_helpers/single_trait_analysis.jl
# ----------------------------------------------------------------------------------------
def quantile_type7(x: np.ndarray, probs) -> np.ndarray:
    """Julia's `quantile(x, p)` default (alpha = beta = 1, Hyndman-Fan type 7) as Statistics.jl evaluates it:
    h = (n-1) p, a = x_sorted[floor h], b = the next one, result a + (h - floor h) (b - a), each operation rounded
    on its own (numpy's `quantile` switches to b - (b-a)(1-t) for t >= 0.5 and can differ in the last bit)."""
    xs = np.sort(np.asarray(x, dtype=np.float64))
    n = xs.shape[0]
    out = []
    for pr in np.atleast_1d(np.asarray(probs, dtype=np.float64)):
        h = min(max((n - 1) * float(pr), 0.0), float(n - 1))
        lo = int(math.floor(h))
        hi = min(lo + 1, n - 1)
        g = h - lo
        out.append(float(xs[lo]) + g * (float(xs[hi]) - float(xs[lo])))
    return np.array(out)


def get_thresholds(L: np.ndarray, signif_level: Sequence[float]):
    """src/analysis_helpers/single_trait_analysis.jl:13-23: quantile of the per-permutation maxima at 1 - level."""
    peaks = np.max(L, axis=0)
    probs = 1.0 - np.asarray(signif_level, dtype=np.float64)
    return {"probs": probs, "thrs": quantile_type7(peaks, probs)}


# ----------------------------------------------------------------------------------------
# readData.jl (the delimited-text readers either side of the scan path)
# ----------------------------------------------------------------------------------------
def _readdlm(file: str, skipstart: int):
    rows = []
    with open(file) as f:
        for i, line in enumerate(f):
            if i < skipstart or not line.strip():
                continue
            rows.append(line.rstrip("\r\n").split(","))
    return rows


def read_bxd_pheno(file: str) -> np.ndarray:
    """src/readData.jl:159-161: readdlm(file, ','; skipstart=1)[:, 2:end-1]."""
    return np.array([[float(x) for x in r[1:-1]] for r in _readdlm(file, 1)], dtype=np.float64)


def read_bxd_geno(file: str, skipstart: int = 1) -> np.ndarray:
    """src/readData.jl:163-165: readdlm(file, ','; skipstart)[:, 2:2:end]."""
    return np.array([[float(x) for x in r[1::2]] for r in _readdlm(file, skipstart)], dtype=np.float64)


def read_genoprob_exclude_complements(file: str) -> np.ndarray:
    """src/readData.jl:85-96 with getmarkernames = getids = true: numeric block after header/ids, odd columns."""
    return np.array([[float(x) for x in r[1:][0::2]] for r in _readdlm(file, 1)], dtype=np.float64)
