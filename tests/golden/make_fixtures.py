"""Generates the committed fixtures (run in the build container, where /root/reference exists):

  bxd_kinship.npy  : the 79 x 79 BXD kinship the reference ships
                     (/root/reference/test/run-lmmlite_R/processed_bxdData/BXDkinship.csv) — the only
                     data input of the reference's tests that survives in this checkout
                     (genotypes/phenotypes are listed in .MISSING_LARGE_BLOBS).
  oracle_small.npz : seeded synthetic genotypes/traits on that kinship and the oracle's outputs,
                     a regression pin for the oracle and a GPU parity case on a realistic K.
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
