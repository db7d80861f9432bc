"""Generates the committed fixtures (run in the build container, where /root/reference exists):

  bxd_kinship.npy  : the 79 x 79 BXD kinship the reference ships
                     (/root/reference/test/run-lmmlite_R/processed_bxdData/BXDkinship.csv) — the only
                     data input of the reference's tests that survives in this checkout
                     (genotypes/phenotypes are listed in .MISSING_LARGE_BLOBS).
  oracle_small.npz : seeded synthetic genotypes/traits on that kinship and the oracle's outputs,
                     a regression pin for the oracle and a GPU parity case on a realistic K.
  oracle_scan.npz  : the single-trait paths on the same kinship: scan (null, REML), scan with permutations (the
                     0-based shuffle indices are part of the fixture), scan assumption="alt", bulkscan_null
                     (per-trait Brent), lod2log10p and get_thresholds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
import blmm_oracle as orc  # noqa: E402
from blmm_b200 import synth  # noqa: E402

src = "/root/reference/test/run-lmmlite_R/processed_bxdData/BXDkinship.csv"
rows = [l.strip().split(",") for l in open(src)]
try:
    K = np.array(rows, dtype=np.float64)
except ValueError:  # header row / row names
    K = np.array([r[1:] if not _isnum(r[0]) else r for r in rows[1:]], dtype=np.float64)
assert K.shape == (79, 79), K.shape
np.save(os.path.join(HERE, "bxd_kinship.npy"), K)

G = synth.make_geno(79, 150, seed=5)
Y = synth.make_pheno(G, K, 40, seed=6)
grid = np.arange(10) / 10.0
r = orc.bulkscan_null_grid(Y, G, K, grid)
a = orc.bulkscan_alt_grid(Y, G, K, grid)
np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), Y=Y, G=G, null_L=r.L, null_h2=r.h2_null_list,
                    alt_L=a.L, alt_h2_panel=a.h2_panel)
print("fixtures written", K.shape, K.min(), K.max())

# ---- single-trait paths --------------------------------------------------------------------------
Ut, lam = orc.decompose(K)
y = Y[:, 3:4]
Cv = synth.make_covar(79)[:, :1]
s_null = orc.scan(y, G, K, covar=Cv, reml=True, Ut=Ut, lam=lam)
perm_idx = orc.make_perm_indices(79, 64, 11)
s_perm = orc.scan(y, G, K, permutation_test=True, perm_idx=perm_idx, Ut=Ut, lam=lam)
s_alt = orc.scan(y, G[:, :60], K, assumption="alt", prior_variance=float(np.var(y, ddof=1)), prior_sample_size=0.1,
                 Ut=Ut, lam=lam)
b_null = orc.bulkscan_null(Y[:, :8], G, K, reml=True, prior_variance=0.0, Ut=Ut, lam=lam)
thr = orc.get_thresholds(s_perm["L_perms"], [0.1, 0.05])
np.savez_compressed(os.path.join(HERE, "oracle_scan.npz"), Ut=Ut, lam=lam, covar=Cv,
                    null_lod=s_null["lod"], null_h2=s_null["h2_null"], null_sigma2=s_null["sigma2_e"],
                    perm_idx=perm_idx, perm_lod=s_perm["lod"], perm_L=s_perm["L_perms"], perm_h2=s_perm["h2_null"],
                    alt_lod=s_alt["lod"], alt_h2_each=s_alt["h2_each_marker"], alt_h2_null=s_alt["h2_null"],
                    bnull_L=b_null.L, bnull_h2=b_null.h2_null_list,
                    log10p=orc.lod2log10p(s_null["lod"], 1), thr=thr["thrs"])
print("scan fixtures written")
