"""DRY RUN ONLY — not a pin, never committed: writes oracle-made stand-ins for tests/golden/ref_outputs/ into a
scratch directory (default /tmp/fake_ref), so that the code paths of tests/test_reference_fixtures.py can be exercised
(BLMM_REF_OUTPUTS=/tmp/fake_ref pytest tests/test_reference_fixtures.py) before anyone has run the real generator,
tests/golden/make_reference_fixtures.jl, under Julia.  Same file names and shapes as that script writes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import blmm_oracle as orc

INP = os.path.join(ROOT, "tests", "golden", "ref_inputs")
OUT = sys.argv[1] if len(sys.argv) > 1 else "/tmp/fake_ref"
assert "golden" not in OUT, "stand-ins must not be written next to the real fixtures"
os.makedirs(OUT, exist_ok=True)
ld = lambda n: np.loadtxt(f"{INP}/{n}.csv", delimiter=",", ndmin=2)
G, K, Y, Z, w = ld("G"), ld("K"), ld("Y"), ld("Covar"), ld("weights")[:, 0]


def wr(n, a):
    np.savetxt(f"{OUT}/{n}.csv", np.atleast_1d(a), delimiter=",", fmt="%.17g")


Ut, lam = orc.decompose(K)
wr("eig_U", Ut.T); wr("eig_lambda", lam); wr("kinship", orc.calc_kinship(G))
grid = np.arange(10) / 10
r = orc.bulkscan_null_grid(Y, G, K, grid, Ut=Ut, lam=lam); wr("nullgrid_L", r.L); wr("nullgrid_h2", r.h2_null_list)
r = orc.bulkscan_null_grid(Y, G, K, grid, Covar=Z, reml=True, Ut=Ut, lam=lam)
wr("nullgrid_cov_reml_L", r.L); wr("nullgrid_cov_reml_h2", r.h2_null_list)
r = orc.bulkscan_null_grid(Y, G, K, grid, weights=w); wr("nullgrid_weights_L", r.L)
for tag, reml in (("altgrid", False), ("altgrid_reml", True)):
    a = orc.bulkscan_alt_grid(Y, G, K, grid, reml=reml, Ut=Ut, lam=lam); wr(tag + "_L", a.L); wr(tag + "_h2panel", a.h2_panel)
r = orc.bulkscan_null(Y, G, K, reml=True, prior_variance=0.0, Ut=Ut, lam=lam)
wr("nullexact_reml_L", r.L); wr("nullexact_reml_h2", r.h2_null_list)
r = orc.bulkscan_null(Y, G, K, Covar=Z, optim_interval=4, Ut=Ut, lam=lam)
wr("nullexact_cov_oi4_L", r.L); wr("nullexact_cov_oi4_h2", r.h2_null_list)
y = Y[:, 2:3]
for tag, reml in (("ml", False), ("reml", True)):
    s = orc.scan(y, G, K, reml=reml, Ut=Ut, lam=lam)
    wr(f"scan_null_{tag}_lod", s["lod"]); wr(f"scan_null_{tag}_scalars", [s["sigma2_e"], s["h2_null"]])
perm = orc.make_perm_indices(79, 64, 0); wr("perms_idx0", perm)
s = orc.scan(y, G, K, permutation_test=True, perm_idx=perm, Ut=Ut, lam=lam)
wr("perms_L", s["L_perms"]); wr("perms_lod", s["lod"]); wr("perms_scalars", [s["sigma2_e"], s["h2_null"]])
wr("perms_thresholds", orc.get_thresholds(s["L_perms"], [0.1, 0.05])["thrs"])
lod = orc.scan(y, G, K, Ut=Ut, lam=lam)["lod"]
wr("lod2log10p_df1", orc.lod2log10p(lod, 1)); wr("lod2log10p_df3", orc.lod2log10p(lod, 3))
a = orc.scan(y, G, K, assumption="alt", Ut=Ut, lam=lam); wr("scan_alt_lod", a["lod"])
print("stand-ins (oracle-made, NOT the reference) in", OUT)
