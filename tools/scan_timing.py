"""Development aid: where the scan kernel's warps spend their time.

    python tools/scan_timing.py build      (here: second library with -DBLMM_SCAN_TIMING under tools/_timing/)
    python tools/scan_timing.py run        (GPU box: BXD-shape alt-grid / null-grid, prints per-phase cycle shares)
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tools", "_timing")
VAR = os.environ.get("TIMING_VARIANT", "")      # e.g. "EARLY_GRANT,TESTWAIT" -> -DBLMM_PP_EARLY_GRANT -DBLMM_PP_TESTWAIT
LIB = os.path.join(OUT, f"libblmm_b200_timing{('_' + VAR.replace(',', '_')) if VAR else ''}.so")
CSRC = os.path.join(ROOT, "bulklmm.jl_b200", "csrc")
PHASES = ["wait stage data (full)", "wait turn (ping-pong)", "DMMA loop", "release + per-k epilogue", "final epilogue + stores",
          "trait tile switch", "-", "-"]


def build():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "bulklmm.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    objs = []
    procs = []
    for src in b.SOURCES:
        o = os.path.join(OUT, src.replace(".cu", (VAR.replace(",", "_") + ".o")))
        objs.append(o)
        procs.append(subprocess.Popen([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")] + b.NVCC_FLAGS + ["-DBLMM_SCAN_TIMING"] + [f"-DBLMM_PP_{v}" for v in VAR.split(",") if v] + ["-c", os.path.join(CSRC, src), "-o", o]))
    assert all(p.wait() == 0 for p in procs)
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-shared", "-o", LIB] + objs + ["-lcusolver", "-Xlinker", "-rpath,/usr/local/cuda/lib64"], check=True)
    print(LIB)


def run():
    os.environ["BLMM_B200_LIB"] = LIB
    sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
    import ctypes as C
    import numpy as np, torch
    from blmm_b200 import Engine, synth, _lib as L
    m = int(sys.argv[2]) if len(sys.argv) > 2 else synth.BXD_M
    n, p = synth.BXD_N, synth.BXD_P
    Y, G, K = synth.make_problem(n, p, m)
    eng = Engine(0)
    lib = L.load()
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
    nb = 148
    buf = np.zeros((nb, 16, 8), dtype=np.int64)
    for name, method, hptr in (("alt-grid", L.METHOD_ALT_GRID, dH.data_ptr()), ("null-grid", L.METHOD_NULL_GRID, dh.data_ptr())):
        o, keep = eng.make_opts(method=method, h2_grid=grid, mem_space=L.MEM_DEVICE)
        for rep in range(3):
            eng.bulkscan_raw(pr, o, dL.data_ptr(), hptr)
            eng.sync()
        rc = lib.blmm_debug_scan_timing(buf.ctypes.data_as(C.c_void_p), nb)
        assert rc == 0
        tot = buf.sum(axis=2)
        print(f"== {name}: scan kernel {eng.last_scan_ms():.3f} ms (instrumented build); warp-cycles mean {tot.mean():.3e} min {tot.min():.3e} max {tot.max():.3e}")
        share = buf.sum(axis=(0, 1)) / buf.sum()
        for i, ph in enumerate(PHASES[:6]):
            g0 = buf[:, :8, i].sum() / buf[:, :8, :].sum()
            g1 = buf[:, 8:, i].sum() / buf[:, 8:, :].sum()
            print(f"   {ph:28s} {100*share[i]:6.2f}%   group0 {100*g0:6.2f}%  group1 {100*g1:6.2f}%")
        tr = np.zeros((16, 64, 4), dtype=np.int64)
        assert lib.blmm_debug_scan_trace(tr.ctypes.data_as(C.c_void_p)) == 0
        t0 = tr[:, 0, 0].min()
        print("   trace CTA 0, iterations 64..: per warp (sub-partition 0: warps 0,4 = group 0; 8,12 = group 1): full-ok, turn-ok, dmma-end, epi-end (cycles)")
        for itx in range(0, 6):
            for w in (0, 4, 8, 12):
                print(f"     it {itx+64} warp {w:2d}: " + " ".join(f"{int(x - t0):8d}" for x in tr[w, itx]))
        per_sm = buf.sum(axis=(1, 2)) / 16
        print("   per-CTA mean warp cycles: min %.3e max %.3e" % (per_sm.min(), per_sm.max()))


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
