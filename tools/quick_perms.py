"""Ad-hoc device-resident timing of configs[3] (1 trait x 10 000 permutations x 7 321 markers); development aid, also the
command whose ncu launch list (profiles/launches_perms_r02.csv) shows where a permutation step's time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
import numpy as np, torch
from blmm_b200 import Engine, synth, _lib as L

def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n, p, nperms = synth.BXD_N, synth.BXD_P, 10000
    G = synth.make_geno(n, p, seed=p)
    K = synth.calc_kinship_host(G)
    y = synth.make_pheno(G, K, 1112, seed=35554)[:, 1111:1112]
    eng = Engine(0)
    U, lam, _ = eng.decompose(K)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a).T)).to(dev)
    dY, dG, dC, dU, dl = t(y), t(G), t(np.ones((n, 1))), t(U), torch.from_numpy(lam).to(dev)
    dperm = torch.from_numpy(np.ascontiguousarray(synth.make_perm_indices(n, nperms, 0).T.astype(np.int32))).to(dev)
    lod = torch.empty(p, dtype=torch.float64, device=dev)
    Lp = torch.empty((nperms, p), dtype=torch.float64, device=dev)
    mx = torch.empty(nperms, dtype=torch.float64, device=dev)
    sc = torch.empty(2, dtype=torch.float64, device=dev)
    pr = eng.make_problem(n, p, 1, 1, dY.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    o, _ = eng.make_opts(prior_variance=0.0, mem_space=L.MEM_DEVICE)
    eng.set_profiling(True)
    st = torch.cuda.ExternalStream(eng.stream, device=dev)
    for rep in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(st)
        eng.scan_perms_raw(pr, o, dperm.data_ptr(), nperms, lod.data_ptr(), Lp.data_ptr(), mx.data_ptr(), sc.data_ptr(), sc.data_ptr() + 8)
        b.record(st)
        eng.sync()
        torch.cuda.synchronize()
        print(f"perms: step {a.elapsed_time(b):.3f} ms, scan kernel {eng.last_scan_ms():.3f} ms", flush=True)

main()
