import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import blmm_oracle as orc
from blmm_b200 import Engine, scan, synth
eng = Engine(0)
def ref_at(h2, y, G, K, Ut, lam, perm, n):
    y0, X0, l0 = orc.transform_rotation(y, G, K, Ut=Ut, lam=lam)
    w = orc.make_weights(h2, l0)
    est = orc.wls(y0, X0[:, :1], w, [0.0, 0.0])
    r0 = (y0 - X0[:, :1] @ est.b) * np.sqrt(w)[:, None]
    X00 = orc.resid(X0[:, 1:] * np.sqrt(w)[:, None], X0[:, :1] * np.sqrt(w)[:, None])
    rp = orc.shuffle_vector(r0[:, 0], perm)
    rp = rp / np.linalg.norm(rp, axis=0)
    X00 = X00 / np.linalg.norm(X00, axis=0)
    return orc.r2lod(X00.T @ rp, n)
for n, p, sg, sy, col in ((79, 150, 109, 209, 1), (79, 200, 1, 2, 1), (79, 200, 1, 2, 0), (79, 200, 109, 209, 1), (79,150,1,2,1)):
    Y, G, K = synth.make_problem(n, p, 4, seed_g=sg, seed_y=sy)
    Ut, lam = orc.decompose(K)
    dec = (np.asfortranarray(Ut.T), lam)
    perm = synth.make_perm_indices(n, 130, rndseed=3)
    y = Y[:, col:col+1]
    for rep in range(2):
        s = scan(y, G, K, permutation_test=True, perm_idx=perm, decomposition=dec, engine=eng)
        r = orc.scan(y, G, K, permutation_test=True, perm_idx=perm, Ut=Ut, lam=lam)
        Lr = ref_at(s.h2_null, y, G, K, Ut, lam, perm, n)
        print(n, p, sg, sy, col, "h2", s.h2_null, r["h2_null"], "lod err vs orc", np.abs(s.lod - r["lod"]).max(),
              "vs ref_at", np.abs(s.lod - Lr[:, 0]).max(), "perm vs ref_at", np.abs(s.L_perms - Lr[:, 1:]).max(), flush=True)
