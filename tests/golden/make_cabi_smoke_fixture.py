"""Writes tests/golden/cabi_smoke.bin: inputs and ORACLE outputs of a small alt-grid scan and a permutation scan, as
raw little-endian arrays that tests/cabi_smoke.c reads with fread (no Python on the consuming side).

Layout: int64 n, p, m, nperms, ngrid; then Float64 arrays, column-major:
  Y[n*m] G[n*p] U[n*n] lambda[n] grid[ngrid] | alt_L[p*m] alt_h2panel[p*m] | y1[n] perm_idx(int32)[n*nperms]
  perm_lod[p] perm_L[p*nperms] perm_h2 perm_sigma2"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import blmm_oracle as orc  # noqa: E402
from blmm_b200 import synth  # noqa: E402

n, p, m, nperms = 79, 64, 40, 33
Y, G, K = synth.make_problem(n, p, m, seed_g=41, seed_y=42)
Ut, lam = orc.decompose(K)
grid = np.arange(10) / 10.0
alt = orc.bulkscan_alt_grid(Y, G, K, grid, Ut=Ut, lam=lam)
perm = synth.make_perm_indices(n, nperms, 5).astype(np.int32)
y1 = Y[:, 7:8]
sc = orc.scan(y1, G, K, permutation_test=True, perm_idx=perm, Ut=Ut, lam=lam)
F = lambda a: np.asfortranarray(np.asarray(a, dtype=np.float64)).tobytes(order="F")
with open(os.path.join(HERE, "cabi_smoke.bin"), "wb") as f:
    f.write(np.array([n, p, m, nperms, len(grid)], dtype=np.int64).tobytes())
    for a in (Y, G, Ut.T, lam, grid, alt.L, alt.h2_panel, y1):
        f.write(F(a))
    f.write(np.asfortranarray(perm).tobytes(order="F"))
    for a in (sc["lod"], sc["L_perms"], [sc["h2_null"]], [sc["sigma2_e"]]):
        f.write(F(a))
print("wrote cabi_smoke.bin", os.path.getsize(os.path.join(HERE, "cabi_smoke.bin")), "bytes")
