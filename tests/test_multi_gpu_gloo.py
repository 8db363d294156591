"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU): disjoint shard ranges + the end-of-run reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from riichienv_b200.multi_gpu import RunStats, reduce_stats, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(3, world, rank, 1000)
    st = RunStats(elapsed_ms=10.0 + rank, kernel_ms=9.0 + rank, e2e_s=1.0 + rank, env_steps=100.0 * (rank + 1),
                  e2e_steps=50.0, games=hi - lo, score_sum=1.0)
    red = reduce_stats(st, dist, torch, "cpu")
    out[rank] = (lo, hi, red.elapsed_ms, red.env_steps, red.games, red.e2e_s)
    dist.destroy_process_group()


def test_gloo_world2_shards_and_reduction():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (lo0, hi0, t0, s0, g0, e0), (lo1, hi1, t1, s1, g1, e1) = out[0], out[1]
    assert (lo0, hi0) == (6000, 7000) and (lo1, hi1) == (7000, 8000)       # disjoint, contiguous
    assert t0 == t1 == 11.0 and e0 == e1 == 2.0                            # max over ranks
    assert s0 == s1 == 300.0 and g0 == g1 == 2000.0                        # sums


def test_shards_cover_without_overlap():
    seen = set()
    for k in range(3):
        for r in range(8):
            lo, hi = shard_range(k, 8, r, 125000)
            assert not (set(range(lo, hi, 12500)) & seen)
            seen |= set(range(lo, hi, 12500))
    assert len(seen) == 3 * 8 * 10
