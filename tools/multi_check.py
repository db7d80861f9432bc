"""One process, N GPUs behind ONE context (blmm_create_multi): BXD-shape timings of the in-library multi-GPU path.

    python tools/multi_check.py [--gpus N] [--steps K]

Prints one JSON line with, for alt-grid / null-grid / permutations at BASELINE.json's shapes:
  * host      : one blocking C-ABI call on the whole problem, pageable and pinned host buffers (ms per call)
  * resident  : device pointers on the primary GPU; NCCL broadcast/scatter of the inputs, sharded scans, NCCL gather
                of the result slabs into the primary's arrays: ms per call (CUDA events on the primary's stream) and
                the gather's own device time (blmm_last_gather_ms)
and checks every multi-GPU result bit for bit against the one-GPU context."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
import numpy as np
import torch

GRID = np.arange(10) / 10.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=torch.cuda.device_count())
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--m", type=int, default=35554)
    args = ap.parse_args()
    from blmm_b200 import Engine, synth, _lib as L
    nd = args.gpus
    n, p, m = 79, 7321, args.m
    G = synth.make_geno(n, p, seed=p)
    K = synth.calc_kinship_host(G)
    Y = synth.make_pheno(G, K, m, seed=35554)
    one = Engine(0)
    many = Engine(devices=list(range(nd)))
    U, lam, _ = one.decompose(K)
    Cv = np.ones((n, 1))
    dev0 = torch.device("cuda:0")

    def cm(a):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).T))

    ins_h = [cm(Y), cm(G), cm(Cv), cm(U), torch.from_numpy(lam.copy())]
    ins_d = [t.to(dev0) for t in ins_h]
    out = {"gpus": nd, "n": n, "p": p, "m": m, "steps": args.steps, "workloads": {}}
    # the ceiling of any host-buffer call at N GPUs: pinned device-to-host copies from all GPUs at once
    nb = 1 << 29
    hb = [torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(nd)]
    db = [torch.empty(nb, dtype=torch.uint8, device=f"cuda:{d}") for d in range(nd)]
    best = 1e9
    for _ in range(3):
        for d in range(nd):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d in range(nd):
            hb[d].copy_(db[d], non_blocking=True)
        for d in range(nd):
            torch.cuda.synchronize(d)
        best = min(best, time.perf_counter() - t0)
    out["host_ingest_gbs_all_gpus_at_once"] = nd * nb / best / 1e9
    del hb, db
    stream = torch.cuda.ExternalStream(many.stream, device=dev0)

    for name, method in (("alt-grid", L.METHOD_ALT_GRID), ("null-grid", L.METHOD_NULL_GRID)):
        alt = name == "alt-grid"
        shapes = [(m, p), (m, p) if alt else (m,)]
        res = {}
        # reference result: one GPU, device resident
        ref = [torch.empty(s, dtype=torch.float64, device=dev0) for s in shapes]
        o_d, k1 = one.make_opts(method=method, h2_grid=GRID, mem_space=L.MEM_DEVICE)
        pr_d = one.make_problem(n, p, m, 1, *[t.data_ptr() for t in ins_d])
        one.bulkscan_raw(pr_d, o_d, ref[0].data_ptr(), ref[1].data_ptr())
        one.sync()
        ref_h = [t.cpu() for t in ref]
        # host buffers through the multi-GPU context
        o_h, k2 = many.make_opts(method=method, h2_grid=GRID, mem_space=L.MEM_HOST)
        for kind in ("pageable", "pinned"):
            ins = [t.clone() for t in ins_h]
            outs = [torch.zeros(s, dtype=torch.float64) for s in shapes]
            if kind == "pinned":
                ins = [t.pin_memory() for t in ins]
                outs = [t.pin_memory() for t in outs]
            pr = many.make_problem(n, p, m, 1, *[t.data_ptr() for t in ins])
            many.bulkscan_raw(pr, o_h, outs[0].data_ptr(), outs[1].data_ptr())
            per = []
            for _ in range(args.steps):
                t0 = time.perf_counter()
                many.bulkscan_raw(pr, o_h, outs[0].data_ptr(), outs[1].data_ptr())
                per.append((time.perf_counter() - t0) * 1e3)
            ok = all(torch.equal(a, b) for a, b in zip(outs, ref_h))
            res["host_" + kind] = {"ms": float(np.mean(per)), "each": [round(x, 2) for x in per], "bit_equal": ok}
        # device resident through the multi-GPU context (NCCL)
        outs_d = [torch.zeros(s, dtype=torch.float64, device=dev0) for s in shapes]
        o_m, k3 = many.make_opts(method=method, h2_grid=GRID, mem_space=L.MEM_DEVICE)
        pr_m = many.make_problem(n, p, m, 1, *[t.data_ptr() for t in ins_d])
        torch.cuda.synchronize()
        for _ in range(2):
            many.bulkscan_raw(pr_m, o_m, outs_d[0].data_ptr(), outs_d[1].data_ptr())
            many.sync()
        ms, gms = [], []
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            many.bulkscan_raw(pr_m, o_m, outs_d[0].data_ptr(), outs_d[1].data_ptr())
            b.record(stream)
            many.sync()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            gms.append(many.last_gather_ms())
        ok = all(torch.equal(a.cpu(), b) for a, b in zip(outs_d, ref_h))
        res["resident_nccl"] = {"ms": float(np.mean(ms)), "gather_ms": float(np.mean(gms)), "bit_equal": ok,
                                "gathered_bytes": int(sum(int(np.prod(s)) for s in shapes) * 8 * (nd - 1) / nd)}
        # one GPU, device resident, same clock
        ms1 = []
        s1 = torch.cuda.ExternalStream(one.stream, device=dev0)
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s1)
            one.bulkscan_raw(pr_d, o_d, ref[0].data_ptr(), ref[1].data_ptr())
            b.record(s1)
            one.sync()
            ms1.append(a.elapsed_time(b))
        res["resident_one_gpu_ms"] = float(np.mean(ms1))
        out["workloads"][name] = res
        del ref, outs_d

    # permutations: configs[3]
    nperms = 10000
    idx = synth.make_perm_indices(n, nperms, 0)
    y = synth.make_pheno(G, K, 1112, seed=35554)[:, 1111:1112]
    dperm = torch.from_numpy(np.ascontiguousarray(idx.T.astype(np.int32))).to(dev0)
    iny = [cm(y).to(dev0)] + ins_d[1:]
    o_p, _ = many.make_opts(prior_variance=0.0, mem_space=L.MEM_DEVICE)
    o_1, _ = one.make_opts(prior_variance=0.0, mem_space=L.MEM_DEVICE)
    res = {}
    bufs = {}
    for nm, E in (("one", one), ("many", many)):
        lod = torch.empty(p, dtype=torch.float64, device=dev0)
        Lp = torch.empty((nperms, p), dtype=torch.float64, device=dev0)
        mx = torch.empty(nperms, dtype=torch.float64, device=dev0)
        sc = torch.empty(2, dtype=torch.float64, device=dev0)
        pr = E.make_problem(n, p, 1, 1, *[t.data_ptr() for t in iny])
        o = o_p if nm == "many" else o_1
        st = torch.cuda.ExternalStream(E.stream, device=dev0)
        ms, gms = [], []
        for i in range(args.steps + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            E.scan_perms_raw(pr, o, dperm.data_ptr(), nperms, lod.data_ptr(), Lp.data_ptr(), mx.data_ptr(), sc.data_ptr(),
                             sc.data_ptr() + 8)
            b.record(st)
            E.sync()
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(a.elapsed_time(b))
                gms.append(E.last_gather_ms())
        bufs[nm] = (lod.cpu(), Lp.cpu(), mx.cpu())
        res[nm] = {"ms": float(np.mean(ms)), "gather_ms": float(np.mean(gms))}
    res["bit_equal"] = all(torch.equal(a, b) for a, b in zip(bufs["one"], bufs["many"]))
    out["workloads"]["perms"] = res
    print(json.dumps(out), flush=True)
    one.close()
    many.close()
    ok = all(v.get("bit_equal", True) for w in out["workloads"].values() for v in (w.values() if isinstance(w, dict) else [])
             if isinstance(v, dict)) and out["workloads"]["perms"]["bit_equal"]
    sys.exit(0 if ok else 1)


main()
