"""Host-side multi-GPU logic on CPU: world_size 2 over gloo (the N > 1 path of bench.py and dist.py)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from blmm_b200 import dist as bd
    j0, j1 = bd.shard_range(m, world, rank)
    full = np.arange(5 * m, dtype=np.float64).reshape(5, m)
    got = bd.gather_columns(full[:, j0:j1], m, axis=1)
    vec = bd.gather_columns(np.arange(j0, j1, dtype=np.float64), m)
    q.put((rank, bool(np.array_equal(got, full)), bool(np.array_equal(vec, np.arange(m)))))
    dist.destroy_process_group()


def test_shard_ranges_cover_exactly():
    from blmm_b200 import dist as bd
    for m in (1, 7, 35554, 10001):
        for world in (1, 2, 4, 8):
            r = bd.all_ranges(m, world)
            assert r[0][0] == 0 and r[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_gather_columns_world2_gloo():
    import sys
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    m = 37  # ragged: 18 + 19 columns
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(ok1 and ok2 for _, ok1, ok2 in res), res
