"""Multi-GPU plumbing: one process per GPU (torchrun), traits (bulkscan) or permutations (scan) sharded
by contiguous column blocks, G / (U, lambda) replicated, no collective on the data path.  NCCL (gloo in
the CPU tests) is used only to gather the small per-column result vectors (h2_null_list, per-permutation
max LOD) and, on request, the LOD slabs.

The reference has no distributed layer (README.md:66-72 only mentions that trait blocks are
independent); the column split mirrors its own `nb` trait blocks (src/bulkscan.jl:263-309)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(m: int, world: int, rank: int) -> Tuple[int, int]:
    """Columns [j0, j1) of rank `rank`: contiguous, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return rank * m // world, (rank + 1) * m // world


def all_ranges(m: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(m, world, r) for r in range(world)]


def gather_columns(local: np.ndarray, m: int, axis: int = -1):
    """All-gather column shards of unequal width into the full array on every rank.
    `local` holds this rank's columns along `axis`.  Works on any torch.distributed backend.
    Meant for the small per-column vectors (h2_null_list, per-permutation maxima): the shards are host arrays and take a
    round trip through the device.  For a whole LOD matrix use ONE multi-GPU context instead (`Engine(devices=[...])`,
    blmm_create_multi): with host buffers every GPU writes its slab of the caller's array directly, and with device
    pointers the slabs are gathered by NCCL without touching the host."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    ranges = all_ranges(m, world)
    width = max(j1 - j0 for j0, j1 in ranges)
    loc = np.moveaxis(np.asarray(local), axis, 0)
    assert loc.shape[0] == ranges[rank][1] - ranges[rank][0], "local shard width does not match shard_range"
    pad = np.zeros((width,) + loc.shape[1:], dtype=loc.dtype)
    pad[: loc.shape[0]] = loc
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy()[: j1 - j0] for o, (j0, j1) in zip(outs, ranges)]
    return np.moveaxis(np.concatenate(parts, axis=0), 0, axis)


def bulkscan_sharded(Y, G, K, engine=None, gather_L: bool = False, **kw):
    """bulkscan with traits sharded over the ranks of the default process group.  Every rank returns
    its own L slab (and the gathered h2_null_list for the null methods); gather_L=True also
    assembles the full p x m LOD matrix on every rank."""
    import torch.distributed as dist
    from .api import bulkscan

    world, rank = dist.get_world_size(), dist.get_rank()
    Y = np.asarray(Y)
    m = Y.shape[1]
    j0, j1 = shard_range(m, world, rank)
    out = bulkscan(Y[:, j0:j1], G, K, engine=engine, **kw)
    out.columns = (j0, j1)
    if hasattr(out, "h2_null_list"):
        out.h2_null_list = gather_columns(out.h2_null_list, m)
    if gather_L:
        out.L = gather_columns(out.L, m, axis=1)
        if hasattr(out, "h2_panel"):
            out.h2_panel = gather_columns(out.h2_panel, m, axis=1)
        out.columns = (0, m)
    return out


def scan_perms_sharded(y, g, K, perm_idx, engine=None, **kw):
    """scan(...; permutation_test=true) with the permutation columns sharded over the ranks; the
    per-permutation maxima (what get_thresholds needs) are gathered on every rank."""
    import torch.distributed as dist
    from .api import scan

    world, rank = dist.get_world_size(), dist.get_rank()
    perm_idx = np.asarray(perm_idx)
    nperms = perm_idx.shape[1]
    j0, j1 = shard_range(nperms, world, rank)
    out = scan(y, g, K, permutation_test=True, perm_idx=perm_idx[:, j0:j1], engine=engine, **kw)
    out.columns = (j0, j1)
    out.max_lod = gather_columns(out.max_lod, nperms)
    return out
