#!/usr/bin/env python
"""bench.py — marker x trait LOD tests/sec on the BXD-shape bulkscan (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload alt-grid|null-grid] [--impl reference]

One "step" = one full bulkscan call (rotation by U' -> per-trait null statistics over the h2 grid ->
weight-folded marker operand -> fused DMMA scan with the LOD / max-over-grid epilogue) on synthetic
BXD-shape data (n=79, p=7321 markers, m=35554 traits, 10-point h2 grid).  The kinship
eigendecomposition is setup (one-off, timed separately and reported as `setup_ms`).

  value      : whole-job tests/s, inputs resident in HBM, outputs (L, h2_panel) left in HBM
  e2e        : the same call through the C-ABI with HOST (pinned) buffers: H2D of Y,G,Covar,U,lambda
               and D2H of the p x m outputs inside the timed region
  roofline   : the fused scan kernel against the measured FP64 tensor peak (profiles/fp64_peak_r01.json)
  cpu_baseline: the CPU oracle (numpy restatement of the reference algorithm), all host cores, on a
               bounded trait sample of the same workload

N > 1 (torchrun, one rank per GPU): traits are sharded across ranks (strong scaling: the BXD problem
is fixed), G/U are replicated, there is no data-path collective; NCCL is used for the barrier and
the max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for sub in ("bulklmm.jl_b200", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import numpy as np  # noqa: E402

N_BXD, P_BXD, M_BXD = 79, 7321, 35554
GRID = np.arange(10) / 10.0  # 0.0:0.1:0.9, src/bulkscan.jl:82
README_REF = {"value": 1.23e8, "what": "reference README.md:336-339: bulkscan null-grid, 2.112 s, 16 Julia threads, "
                                       "48x Xeon Silver 4214 (other hardware; not this metric's alt-grid method)"}


def fp64_peak():
    path = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d["cublas_dgemm_tflops_sustained"], "measured: cuBLAS DGEMM 8192^3 sustained on this pool (profiles/fp64_peak_r01.json); MEASURED_PEAKS.json has no FP64 entry"
    return 37.0, "fallback: B200 datasheet FP64 tensor 37 TFLOP/s (no measured file)"


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = max(mx, float(f[1]))
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle, threaded over trait blocks like the reference (Threads.@threads over nb
# blocks with BLAS pinned to one thread, src/bulkscan.jl:252-286)
# ------------------------------------------------------------------------------------------------
def cpu_scan(workload, Y, G, K, Ut, lam, cores):
    import blmm_oracle as orc
    from concurrent.futures import ThreadPoolExecutor
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    m = Y.shape[1]
    nb = max(1, min(cores * 2, m // 64))
    edges = np.linspace(0, m, nb + 1).astype(int)
    fn = orc.bulkscan_alt_grid if workload == "alt-grid" else orc.bulkscan_null_grid

    def work(i):
        return fn(Y[:, edges[i]:edges[i + 1]], G, K, GRID, Ut=Ut, lam=lam)

    def run():
        with ThreadPoolExecutor(cores) as ex:
            return list(ex.map(work, range(nb)))

    t0 = time.perf_counter()
    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            run()
    else:
        run()
    return time.perf_counter() - t0


def cpu_baseline(workload, sample_m, seed=0):
    import blmm_oracle as orc
    from blmm_b200 import synth
    cores = os.cpu_count() or 1
    G = synth.make_geno(N_BXD, P_BXD)
    K = synth.calc_kinship_host(G)
    Y = synth.make_pheno(G, K, sample_m, seed=35554 + seed)
    Ut, lam = orc.decompose(K)
    dt = cpu_scan(workload, Y, G, K, Ut, lam, cores)
    return {"value": P_BXD * sample_m / dt, "unit": "tests/s", "cores": cores, "kind": "port",
            "sample": f"{workload}, all {P_BXD} markers x first {sample_m} of {M_BXD} synthetic BXD-shape traits, "
                      f"{len(GRID)}-point grid, {dt:.2f} s; numpy oracle threaded over trait blocks "
                      f"(reference cannot run: no Julia in the image)",
            "seconds": dt}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (Julia is
    not installed, so oracle/_ref does not exist), all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_m = 2048 if args.workload == "alt-grid" else 8192
    times = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(args.workload, sample_m, seed=i)
        if i >= args.warmup:
            times.append(last["seconds"])
    dt = float(np.mean(times))
    value = P_BXD * sample_m / dt
    line = {"impl": "reference", "metric": "marker x trait LOD tests/sec, BXD-shape bulkscan", "value": value,
            "unit": "tests/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, args.gpus),
            "cpu_baseline": {"value": value, "unit": "tests/s", "cores": last["cores"], "kind": "port",
                             "sample": last["sample"]},
            "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(workload, gpus):
    return {"workload": f"bulkscan {workload}, BXD shape n={N_BXD} p={P_BXD} m={M_BXD}, h2 grid 0:0.1:0.9, ML, "
                        f"c=1 (BASELINE.json configs[{2 if workload == 'alt-grid' else 1}])",
            "n": N_BXD, "p": P_BXD, "m": M_BXD, "ngrid": len(GRID), "outputs": "L and h2_panel (p x m f64 each)"
            if workload == "alt-grid" else "L (p x m f64) and h2_null_list",
            "sharding": f"traits over {gpus} GPU(s), G/U replicated, no data-path collective",
            "l2": "explicit 512 MB L2 flush between timed steps (each step also writes > 2 GB of output)"}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="alt-grid", choices=["alt-grid", "null-grid"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg")
    ap.add_argument("--no-other", action="store_true", help="skip the short runs of the other configs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from blmm_b200 import Engine, synth, _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)

    # ---- synthetic inputs: same generator and seeds on every rank, then shard traits
    n, p, m = N_BXD, P_BXD, M_BXD
    G = synth.make_geno(n, p)
    K = synth.calc_kinship_host(G)
    Y = synth.make_pheno(G, K, m)
    j0, j1 = rank * m // world, (rank + 1) * m // world
    Ysh = np.asfortranarray(Y[:, j0:j1])
    ml = j1 - j0
    Cv = np.ones((n, 1))

    eng = Engine(local)
    t0 = time.perf_counter()
    U, lam, _ = eng.decompose(K)  # setup: cuSOLVER syevd
    setup_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    eng.decompose(K)
    setup_ms_warm = (time.perf_counter() - t0) * 1e3

    method = L.METHOD_ALT_GRID if args.workload == "alt-grid" else L.METHOD_NULL_GRID
    alt = args.workload == "alt-grid"

    def colmajor(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T))

    hY, hG, hC, hU, hl = colmajor(Ysh), colmajor(G), colmajor(Cv), colmajor(U), torch.from_numpy(lam.copy())
    dY, dG, dC, dU, dl = (t.to(dev) for t in (hY, hG, hC, hU, hl))
    dL = torch.empty((ml, p), dtype=torch.float64, device=dev)
    dH = torch.empty((ml, p), dtype=torch.float64, device=dev) if alt else torch.empty(ml, dtype=torch.float64, device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    pr = eng.make_problem(n, p, ml, 1, dY.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
    opts, keep = eng.make_opts(method=method, h2_grid=GRID, mem_space=L.MEM_DEVICE)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)
    eng.set_profiling(True)

    def step():
        eng.bulkscan_raw(pr, opts, dL.data_ptr(), dH.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        eng.sync()

    for _ in range(args.warmup):
        step()
        eng.sync()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    scan_ms = []
    barrier()
    launches0 = eng.launch_count
    wall0 = time.time()
    for a, b in evs:
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush, outside the step's event pair
        a.record(stream)
        step()
        b.record(stream)
        eng.sync()
        scan_ms.append(eng.last_scan_ms())
    barrier()
    wall1 = time.time()
    launches = eng.launch_count - launches0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    value = p * m * args.steps / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel (the fused scan), measured live with CUDA events
    ngrid = len(GRID)
    flops = 2.0 * n * p * ml * (ngrid if alt else 1)
    scan_avg_ms = float(np.mean(scan_ms))
    peak, peak_src = fp64_peak()
    out_bytes = 8.0 * p * ml * (2 if alt else 1)
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    t_tensor = flops / (peak * 1e12)
    t_hbm = out_bytes / (hbm * 1e9)
    ach = flops / (scan_avg_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "scan_traffic_r01.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    roofline = {"bound": "tensor" if t_tensor >= t_hbm else "hbm", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "kernel": "blmm::scan_kernel (FP64 DMMA.8x8x4, TMA bulk copies)",
                "kernel_ms": scan_avg_ms, "kernel_share_of_step": scan_avg_ms * args.steps / sum(a.elapsed_time(b) for a, b in evs),
                "algorithmic_flops_per_launch": flops, "algorithmic_output_bytes_per_launch": out_bytes,
                "t_roof_ms": max(t_tensor, t_hbm) * 1e3, "frac_of_t_roof": max(t_tensor, t_hbm) * 1e3 / scan_avg_ms,
                "peak_source": peak_src, "hbm_gbs": hbm}

    # ---- e2e: host (pinned) buffers through the C-ABI, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.pin_memory()
        pY, pG, pC, pU, pl = pin(hY), pin(hG), pin(hC), pin(hU), pin(hl)
        pL = torch.empty((ml, p), dtype=torch.float64).pin_memory()
        pH = (torch.empty((ml, p), dtype=torch.float64) if alt else torch.empty(ml, dtype=torch.float64)).pin_memory()
        hpr = eng.make_problem(n, p, ml, 1, pY.data_ptr(), pG.data_ptr(), pC.data_ptr(), pU.data_ptr(), pl.data_ptr())
        hopts, keep2 = eng.make_opts(method=method, h2_grid=GRID, mem_space=L.MEM_HOST)
        e2e_steps = max(2, min(args.steps, 5))
        eng.bulkscan_raw(hpr, hopts, pL.data_ptr(), pH.data_ptr())  # warm (allocates staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.bulkscan_raw(hpr, hopts, pL.data_ptr(), pH.data_ptr())  # blocking: returns with results on the host
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        h2d = sum(t.numel() * t.element_size() for t in (pY, pG, pC, pU, pl))
        d2h = pL.numel() * 8 + pH.numel() * 8
        e2e = {"value": p * m * e2e_steps / dt, "unit": "tests/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt / e2e_steps * 1e3, "steps": e2e_steps,
               "note": "per rank bytes; pinned host buffers; wall clock around blocking C-ABI calls"}
        # sanity: host and device paths agree
        assert torch.equal(pL, dL.cpu()), "host-buffer result differs from device-resident result"

    # ---- the other BASELINE.json configs on one GPU, device-resident, a few steps each (context for the
    # headline number, not part of it)
    other = None
    if world == 1 and not args.no_other:
        other = {}

        def timed(fn, reps=3):
            fn(); eng.sync()
            ms, ks = [], []
            for _ in range(reps):
                with torch.cuda.stream(stream):
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); fn(); b.record(stream); eng.sync()
                ms.append(a.elapsed_time(b)); ks.append(eng.last_scan_ms())
            return float(np.mean(ms)), float(np.mean(ks))

        dh = torch.empty(ml, dtype=torch.float64, device=dev)
        for name, meth, kw, fl in (("null-grid", L.METHOD_NULL_GRID, dict(h2_grid=GRID), 2.0 * n * p * ml),
                                   ("null-exact (REML Brent per trait)", L.METHOD_NULL_EXACT, dict(reml=True, prior_variance=0.0),
                                    2.0 * n * p * ml * 3)):
            if alt is False and name == "null-grid":
                continue
            o2, keep3 = eng.make_opts(method=meth, mem_space=L.MEM_DEVICE, **kw)
            t_ms, k_ms = timed(lambda: eng.bulkscan_raw(pr, o2, dL.data_ptr(), dh.data_ptr()))
            other[name] = {"ms_per_step": t_ms, "tests_per_s": p * ml / (t_ms * 1e-3), "scan_kernel_ms": k_ms,
                           "scan_kernel_tflops": fl / (k_ms * 1e-3) / 1e12}
        nperms = 10000
        perm = torch.from_numpy(np.ascontiguousarray(synth.make_perm_indices(n, nperms, 0).T)).to(dev)
        dy1 = dY[1111:1112].contiguous()
        pr1 = eng.make_problem(n, p, 1, 1, dy1.data_ptr(), dG.data_ptr(), dC.data_ptr(), dU.data_ptr(), dl.data_ptr())
        o3, _k = eng.make_opts(method=L.METHOD_NULL_EXACT, prior_variance=0.0, mem_space=L.MEM_DEVICE)
        lod1 = torch.empty(p, dtype=torch.float64, device=dev)
        mx = torch.empty(nperms, dtype=torch.float64, device=dev)
        sc = torch.empty(2, dtype=torch.float64, device=dev)
        for name, lp in (("scan 1 trait x 10000 permutations, per-permutation max only", None),
                         ("scan 1 trait x 10000 permutations, L_perms materialised", dL.data_ptr())):
            t_ms, k_ms = timed(lambda: eng.scan_perms_raw(pr1, o3, perm.data_ptr(), nperms, lod1.data_ptr(), lp,
                                                          mx.data_ptr(), sc.data_ptr(), sc.data_ptr() + 8))
            other[name] = {"ms_per_step": t_ms, "tests_per_s": p * (nperms + 1) / (t_ms * 1e-3), "scan_kernel_ms": k_ms,
                           "scan_kernel_tflops": 2.0 * n * p * (nperms + 1) / (k_ms * 1e-3) / 1e12}

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        cpu = cpu_baseline(args.workload, 2048 if alt else 8192)
        cpu.pop("seconds", None)

    if rank == 0:
        line = {"metric": "marker x trait LOD tests/sec, BXD-shape bulkscan", "value": value, "unit": "tests/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": workload_config(args.workload, world), "roofline": roofline,
                "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "other_workloads": other,
                "setup_ms": {"eigendecomposition_first_call": setup_ms, "eigendecomposition_warm": setup_ms_warm},
                "reference_published": README_REF}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
