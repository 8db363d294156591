"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU): disjoint shard ranges + the end-of-run reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from riichienv_b200.multi_gpu import RunStats, reduce_stats, shard_paths, shard_range


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


def _replay_worker(rank, world, port, paths, out):
    import ctypes as C

    from riichienv_b200 import _abi as A
    from riichienv_b200._lib import check, lib

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_paths(paths, world, rank)
    arr = (C.c_char_p * len(mine))(*[p.encode() for p in mine])
    h, failed = C.c_void_p(), C.c_int(0)
    check(lib().rv_replay_from_files(arr, len(mine), 0, A.RULE_DEFAULT_TENHOU, 2, C.byref(h), C.byref(failed)))
    nr, na = C.c_int64(0), C.c_int64(0)
    check(lib().rv_replay_totals(h, 4, C.byref(nr), C.byref(na)))
    lib().rv_replay_free(h)
    t = torch.tensor([len(mine), nr.value, na.value, failed.value], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out[rank] = (mine, nr.value, t.tolist())
    dist.destroy_process_group()


def test_gloo_world2_replay_files_are_sharded_without_overlap(tmp_path):
    """replay ingestion at N > 1: each rank reads its own files with the bulk reader; the only exchange is the sum of the counts"""
    import ctypes as C

    from riichienv_b200 import _abi as A
    from riichienv_b200._lib import check, lib
    from tests.test_replay import REAL_LOG, simulated_log

    paths = []
    for i in range(5):
        p = tmp_path / f"log{i}.jsonl"
        p.write_text(open(REAL_LOG).read() if i == 0 else "\n".join(simulated_log(2, 70 + i)) + "\n")
        paths.append(str(p))
    paths.append(str(tmp_path / "missing.jsonl"))
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    h, failed = C.c_void_p(), C.c_int(0)
    check(lib().rv_replay_from_files(arr, len(paths), 0, A.RULE_DEFAULT_TENHOU, 1, C.byref(h), C.byref(failed)))
    nr, na = C.c_int64(0), C.c_int64(0)
    check(lib().rv_replay_totals(h, 4, C.byref(nr), C.byref(na)))
    lib().rv_replay_free(h)
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_replay_worker, args=(world, _free_port(), paths, out), nprocs=world, join=True)
    assert sorted(out[0][0] + out[1][0]) == sorted(paths) and not set(out[0][0]) & set(out[1][0])
    assert out[0][2] == out[1][2] == [len(paths), nr.value, na.value, 1]
    assert out[0][1] > 0 and out[1][1] > 0 and out[0][1] + out[1][1] == nr.value
