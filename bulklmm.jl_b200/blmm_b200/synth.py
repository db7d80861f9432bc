"""Synthetic BXD-shaped inputs.

The BXD genotype/phenotype CSVs are absent from the reference checkout, so every test and the
benchmark run on data of the same shape and character generated here (numpy only, seeded).
"""
from __future__ import annotations

import numpy as np

BXD_N, BXD_P, BXD_M = 79, 7321, 35554


def make_geno(n: int, p: int, seed: int = 7321, switch: float = 0.01, fuzz: float = 0.02) -> np.ndarray:
    """RIL-like genotype probabilities: per strain a two-state Markov chain along the markers
    (switch probability `switch`), values {0,1}, a fraction `fuzz` of entries replaced by
    U(0,1) "imputed" probabilities.  Monomorphic columns are resampled: the reference throws on a
    zero-norm marker (src/util.jl:69-71)."""
    rng = np.random.default_rng(seed)
    start = rng.integers(0, 2, size=(n, 1))
    flips = rng.random((n, p - 1)) < switch
    state = np.concatenate([start, flips.astype(np.int64)], axis=1).cumsum(axis=1) % 2
    G = state.astype(np.float64)
    mask = rng.random((n, p)) < fuzz
    G[mask] = rng.random(int(mask.sum()))
    bad = np.where(G.std(axis=0) < 1e-8)[0]
    for j in bad:
        G[:, j] = rng.integers(0, 2, size=n)
        G[0, j], G[1, j] = 0.0, 1.0
    return G


def calc_kinship_host(G: np.ndarray, digits: int = 12) -> np.ndarray:
    """K = 2(G-1/2)(G-1/2)'/p + 1/2, diag 1 (src/kinship.jl:4-14), rounded as the reference's
    tests do (test/generate_test_bxdData.jl:14).  Host helper for input generation only."""
    X = G - 0.5
    K = 2.0 * (X @ X.T) / X.shape[1] + 0.5
    np.fill_diagonal(K, 1.0)
    return np.round(K, digits)


def make_pheno(G: np.ndarray, K: np.ndarray, m: int, seed: int = 35554, qtl_frac: float = 0.10,
               mean: float = 11.0, scale: float = 0.5) -> np.ndarray:
    """y_j = mean + scale*(sqrt(h_j) L_K z_j + sqrt(1-h_j) e_j), h_j ~ U(0,0.9); a fraction of the
    traits carries one marker effect so the LOD matrix has a realistic tail."""
    rng = np.random.default_rng(seed)
    n = G.shape[0]
    w, V = np.linalg.eigh(K)
    LK = V * np.sqrt(np.clip(w, 0.0, None))[None, :]
    h = rng.uniform(0.0, 0.9, size=m)
    Z = rng.standard_normal((n, m))
    E = rng.standard_normal((n, m))
    Y = mean + scale * ((LK @ Z) * np.sqrt(h)[None, :] + E * np.sqrt(1.0 - h)[None, :])
    nq = int(round(qtl_frac * m))
    if nq > 0:
        tr = rng.choice(m, size=nq, replace=False)
        mk = rng.integers(0, G.shape[1], size=nq)
        beta = rng.normal(0.0, 0.5, size=nq)
        Y[:, tr] += G[:, mk] * beta[None, :]
    return Y


def make_covar(n: int, seed: int = 3) -> np.ndarray:
    """[Bernoulli(0.5), N(0,1)] background covariates (scaled config C5; intercept added by caller)."""
    rng = np.random.default_rng(seed)
    return np.column_stack([rng.integers(0, 2, size=n).astype(np.float64), rng.standard_normal(n)])


def make_problem(n: int = BXD_N, p: int = BXD_P, m: int = BXD_M, seed_g: int = 7321, seed_y: int = 35554):
    G = make_geno(n, p, seed=seed_g)
    K = calc_kinship_host(G)
    Y = make_pheno(G, K, m, seed=seed_y)
    return Y, G, K


def make_perm_indices(n: int, nperms: int, rndseed: int = 0) -> np.ndarray:
    """Permutation indices (n x nperms int32, 0-based), the host-side stand-in for the reference's
    MersenneTwister(rndseed)+shuffle (src/transform_helpers.jl:98, src/util.jl:175): in production
    the Julia shim draws them with the reference's own RNG and passes them through the C-ABI."""
    rng = np.random.default_rng(rndseed)
    idx = np.tile(np.arange(n, dtype=np.int32)[:, None], (1, nperms))
    return np.ascontiguousarray(rng.permuted(idx, axis=0))
