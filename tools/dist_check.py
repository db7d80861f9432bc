"""Multi-GPU drop-in check (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py

Traits (bulkscan) and permutation columns (scan) are sharded over the ranks; NCCL gathers h2_null_list, the LOD slabs
and the per-permutation maxima.  Every rank checks that the assembled results equal, bit for bit, the unsharded call
on its own GPU, and rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bulklmm.jl_b200"))
import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from blmm_b200 import Engine, bulkscan, scan, synth, get_thresholds, thresholds_from_max
    from blmm_b200 import dist as bd
    eng = Engine(local)
    Y, G, K = synth.make_problem(79, 700, 1003, seed_g=3, seed_y=4)
    U, lam, _ = eng.decompose(K)
    dec = (U, lam)
    grid = np.arange(10) / 10.0
    ok = {}
    for method in ("null-grid", "alt-grid", "null-exact"):
        kw = dict(method=method, h2_grid=grid, decomposition=dec)
        whole = bulkscan(Y, G, K, engine=eng, **kw)
        sh = bd.bulkscan_sharded(Y, G, K, engine=eng, gather_L=True, **kw)
        good = np.array_equal(sh.L, whole.L)
        if method == "alt-grid":
            good &= np.array_equal(sh.h2_panel, whole.h2_panel)
        else:
            good &= np.array_equal(sh.h2_null_list, whole.h2_null_list)
        ok[method] = bool(good)
    idx = synth.make_perm_indices(79, 501, 9)
    whole = scan(Y[:, 5], G, K, permutation_test=True, perm_idx=idx, decomposition=dec, engine=eng)
    sh = bd.scan_perms_sharded(Y[:, 5], G, K, idx, decomposition=dec, engine=eng)
    ok["perms max_lod"] = bool(np.array_equal(sh.max_lod, whole.max_lod))
    ok["perms thresholds"] = bool(np.array_equal(thresholds_from_max(sh.max_lod, [0.1, 0.05], engine=eng).thrs,
                                                 get_thresholds(whole.L_perms, [0.1, 0.05], engine=eng).thrs))
    flags = torch.tensor([int(all(ok.values()))], device=f"cuda:{local}")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "sharded == unsharded, bit for bit", "world": world, "backend": dist.get_backend(),
                          "results": ok, "all_ranks_ok": bool(flags.item())}), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flags.item() else 1)


main()
