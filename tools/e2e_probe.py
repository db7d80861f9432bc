"""Where does the host-buffer (e2e) time of the BXD alt-grid call go?  One process, --gpus N behind one context.

Variants of the same blocking C-ABI call (n=79, p=7321, m=35554, 10-point grid), ms per call:
  pinned_idx      pinned outputs, h2 panel as one-byte indices expanded by the drain threads (default for big panels)
  pinned_f64      pinned outputs, h2 panel as Float64 over PCIe (no host expansion)
  pinned_L_only   pinned output, no h2 panel requested (h2_out = NULL): the PCIe floor of L alone
  pageable_idx    ordinary numpy outputs (ring + drain threads)
  threads_T       pinned_idx with T drain threads
next to the measured rates they must be read against: pinned D2H GB/s per GPU and the host's streaming-store GB/s."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
import numpy as np
import torch

GRID = np.arange(10) / 10.0


def run_variant(nd, steps, pinned, want_h2, m):
    from blmm_b200 import Engine, synth, _lib as L
    n, p = 79, 7321
    G = synth.make_geno(n, p, seed=p)
    K = synth.calc_kinship_host(G)
    Y = synth.make_pheno(G, K, m, seed=35554)
    E = Engine(devices=list(range(nd))) if nd > 1 else Engine(0)
    U, lam, _ = E.decompose(K)

    def cm(a):
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T))
        return t.pin_memory() if pinned else t

    ins = [cm(Y), cm(G), cm(np.ones((n, 1))), cm(U), torch.from_numpy(lam.copy())]
    outs = [torch.zeros((m, p), dtype=torch.float64) for _ in range(2 if want_h2 else 1)]
    if pinned:
        outs = [t.pin_memory() for t in outs]
    pr = E.make_problem(n, p, m, 1, *[t.data_ptr() for t in ins])
    o, keep = E.make_opts(method=L.METHOD_ALT_GRID, h2_grid=GRID, mem_space=L.MEM_HOST)
    E.bulkscan_raw(pr, o, outs[0].data_ptr(), outs[1].data_ptr() if want_h2 else None)
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        E.bulkscan_raw(pr, o, outs[0].data_ptr(), outs[1].data_ptr() if want_h2 else None)
        per.append((time.perf_counter() - t0) * 1e3)
    E.close()
    return {"ms": float(np.mean(per)), "min_ms": float(np.min(per)), "each": [round(x, 2) for x in per]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--m", type=int, default=35554)
    ap.add_argument("--variant", default=None)
    args = ap.parse_args()
    if args.variant:  # child process: environment knobs are read at context creation
        pinned, want_h2 = args.variant.split(",")
        print(json.dumps(run_variant(args.gpus, args.steps, pinned == "1", want_h2 == "1", args.m)))
        return
    from blmm_b200 import _lib as L
    lib = L.load()
    out = {"gpus": args.gpus, "m": args.m, "host_cores": os.cpu_count(), "variants": {}}
    out["host_write_gbs"] = {str(t): lib.blmm_host_write_gbs(t, 1 << 31) for t in (1, 4, 8, 15, 30) if t <= (os.cpu_count() or 1)}
    nb = 1 << 30
    hbuf = torch.empty(nb, dtype=torch.uint8).pin_memory()
    rates = []
    for d in range(args.gpus):
        dbuf = torch.empty(nb, dtype=torch.uint8, device=f"cuda:{d}")
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            hbuf.copy_(dbuf, non_blocking=True)
            torch.cuda.synchronize(d)
            best = min(best, time.perf_counter() - t0)
        rates.append(nb / best / 1e9)
        del dbuf
    out["pinned_d2h_gbs_per_gpu"] = rates
    del hbuf

    # what page-locking the caller's arrays for the duration of a call would cost (the alternative to the ring)
    try:
        rt = torch.cuda.cudart()
        a = np.zeros(1 << 28)  # 2 GiB, touched
        t0 = time.perf_counter()
        r1 = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
        t1 = time.perf_counter()
        r2 = rt.cudaHostUnregister(a.ctypes.data)
        t2 = time.perf_counter()
        out["cudaHostRegister_2GiB_ms"] = {"register": (t1 - t0) * 1e3, "unregister": (t2 - t1) * 1e3, "rc": [int(r1), int(r2)]}
        del a
    except Exception as ex:  # pragma: no cover
        out["cudaHostRegister_2GiB_ms"] = {"error": str(ex)[:200]}

    def child(name, pinned, want_h2, env):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, __file__, "--gpus", str(args.gpus), "--steps", str(args.steps), "--m", str(args.m),
                            "--variant", f"{int(pinned)},{int(want_h2)}"], env=e, capture_output=True, text=True)
        try:
            out["variants"][name] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            out["variants"][name] = {"error": (r.stderr or r.stdout)[-300:]}

    child("pinned_idx", True, True, {"BLMM_B200_H2_TRANSFER": "index"})
    child("pinned_f64", True, True, {"BLMM_B200_H2_TRANSFER": "f64"})
    child("pinned_L_only", True, False, {})
    child("pageable_idx", False, True, {"BLMM_B200_H2_TRANSFER": "index"})
    child("pageable_f64", False, True, {"BLMM_B200_H2_TRANSFER": "f64"})
    if os.environ.get("PROBE_RING_SWEEP"):
        for kb, ns in ((256, 96), (512, 64), (1024, 32), (2048, 24), (4096, 24), (8192, 16)):
            child(f"ring_{kb}KBx{ns}_pageable_idx", False, True,
                  {"BLMM_B200_H2_TRANSFER": "index", "BLMM_B200_RING_SLOT_KB": str(kb), "BLMM_B200_RING_SLOTS": str(ns)})
        for t in (4, 8, 10):
            child(f"threads_{t}_pinned_idx", True, True, {"BLMM_B200_H2_TRANSFER": "index", "BLMM_B200_HOST_THREADS": str(t)})
        for t in (10, 20):
            child(f"threads_{t}_pageable_idx", False, True, {"BLMM_B200_H2_TRANSFER": "index", "BLMM_B200_HOST_THREADS": str(t)})
    for t in (2, 6):
        child(f"threads_{t}_pinned_idx", True, True, {"BLMM_B200_H2_TRANSFER": "index", "BLMM_B200_HOST_THREADS": str(t)})
        child(f"threads_{t}_pageable_idx", False, True, {"BLMM_B200_H2_TRANSFER": "index", "BLMM_B200_HOST_THREADS": str(t)})
    # one traced call of the default path (stderr of the child carries the library's phase times)
    e = dict(os.environ)
    e.update({"BLMM_B200_TRACE": "1"})
    for nm, pin in (("pinned", 1), ("pageable", 0)):
        r = subprocess.run([sys.executable, __file__, "--gpus", str(args.gpus), "--steps", "2", "--m", str(args.m),
                            "--variant", f"{pin},1"], env=e, capture_output=True, text=True)
        out["trace_" + nm] = [ln for ln in r.stderr.splitlines() if "blmm trace" in ln][-max(1, args.gpus):]
    print(json.dumps(out), flush=True)


main()
