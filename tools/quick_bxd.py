"""Ad-hoc device-resident timing of the BXD-shape scans (development aid; bench.py is the contract)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
import numpy as np, torch
from blmm_b200 import Engine, synth, _lib as L

def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else synth.BXD_M
    n, p = synth.BXD_N, synth.BXD_P
    Y, G, K = synth.make_problem(n, p, m)
    eng = Engine(0)
    U, lam, _ = eng.decompose(K)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a).T)).to(dev)
    dY, dG, dC, dU, dl = t(Y), t(G), t(np.ones((n, 1))), t(U), torch.from_numpy(lam).to(dev)
    dL = torch.empty((m, p), dtype=torch.float64, device=dev)
    dH = torch.empty((m, p), dtype=torch.float64, device=dev)
    dh = torch.empty(m, dtype=torch.float64, device=dev)
    pr = eng.make_problem(n, p, m, 1, dY.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    grid = np.arange(10) / 10.0
    eng.set_profiling(True)
    for name, method, hptr in (("alt-grid", L.METHOD_ALT_GRID, dH.data_ptr()), ("alt-grid-noH2", L.METHOD_ALT_GRID, None),
                               ("null-grid", L.METHOD_NULL_GRID, dh.data_ptr())):
        o, keep = eng.make_opts(method=method, h2_grid=grid, mem_space=L.MEM_DEVICE)
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.bulkscan_raw(pr, o, dL.data_ptr(), hptr)
            eng.sync()
            dt = time.perf_counter() - t0
            print(f"{name}: total {dt*1e3:.3f} ms, scan kernel {eng.last_scan_ms():.3f} ms, tests/s {p*m/dt:.3e}", flush=True)
    print("L sample", dL[0, :3].tolist(), "launches", eng.launch_count)

main()
