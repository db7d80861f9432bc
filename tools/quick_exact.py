"""Ad-hoc device-resident timing of bulkscan null-exact (development aid).  usage: quick_exact.py n p m ncov [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
import numpy as np, torch
from blmm_b200 import Engine, synth, _lib as L

def main():
    n, p, m, ncov = (int(x) for x in sys.argv[1:5])
    reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    t0 = time.perf_counter()
    G = synth.make_geno(n, p, seed=p)
    K = synth.calc_kinship_host(G)
    Y = synth.make_pheno(G, K, m, seed=m)
    Cv = np.hstack([np.ones((n, 1))] + ([synth.make_covar(n)[:, :ncov]] if ncov else []))
    c = Cv.shape[1]
    print(f"gen {time.perf_counter()-t0:.1f}s  n={n} p={p} m={m} c={c}", flush=True)
    eng = Engine(0)
    t0 = time.perf_counter(); U, lam, _ = eng.decompose(K); print(f"decompose {time.perf_counter()-t0:.3f}s lam[0]={lam[0]:.4g}", flush=True)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a).T)).to(dev)
    dY, dG, dC, dU, dl = t(Y), t(G), t(Cv), t(U), torch.from_numpy(lam).to(dev)
    dL = torch.empty((m, p), dtype=torch.float64, device=dev)
    dh = torch.empty(m, dtype=torch.float64, device=dev)
    pr = eng.make_problem(n, p, m, c, dY.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    eng.set_profiling(True)
    o, keep = eng.make_opts(method=L.METHOD_NULL_EXACT, reml=True, prior_variance=0.0, mem_space=L.MEM_DEVICE)
    flops = 2.0 * n * p * m * (c + 2)
    for rep in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.bulkscan_raw(pr, o, dL.data_ptr(), dh.data_ptr()); eng.sync()
        dt = time.perf_counter() - t0
        ks = eng.last_scan_ms()
        print(f"null-exact: total {dt*1e3:.3f} ms, scan kernel {ks:.3f} ms ({flops/ks/1e9:.2f} TF/s expanded), tests/s {p*m/dt:.3e}", flush=True)
    h = dh.cpu().numpy()
    print("h2 quantiles", np.quantile(h, [0, .25, .5, .75, 1]), "L sample", dL[0, :3].tolist(), "L max", float(dL.max()), "nan", int(torch.isnan(dL).sum()))

main()
