#!/usr/bin/env python3
"""The whole replay-ingestion pipeline a data loader runs, files -> labelled decision rows.

Input: G hanchan played by this repo's simulator, written as gzip MJAI logs; the file list is repeated REP times (the page
cache holds the files: this times parsing, not the disk).  Timed:
  parse      rv_replay_from_files with 1 / 8 / all host threads (gzip + JSON + KyokuBuilder)          -> files/s, kyoku/s
  pipeline   ReplayBatch.from_files (parse on all threads, flatten, labels, ONE upload) and then, per log position,
             rv_vec_replay_advance + rv_vec_encode (74x34 rows + masks on the device) + labels_of_rows  -> labelled rows/s
usage: time_bulk_load.py [G] [REP]"""
import ctypes as C
import gzip
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from riichienv_b200 import _abi as A  # noqa: E402
from riichienv_b200._lib import check, lib  # noqa: E402
import riichienv_b200.replay as R  # noqa: E402
from tests.test_replay import simulated_log  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REP = int(sys.argv[2]) if len(sys.argv) > 2 else 32
tmp = tempfile.mkdtemp()
paths = []
for seed in range(G):
    p = os.path.join(tmp, f"g{seed}.jsonl.gz")
    with gzip.open(p, "wt") as f:
        f.write("\n".join(simulated_log(2, 800 + seed)) + "\n")
    paths.append(p)
files = paths * REP
L = lib()
arr = (C.c_char_p * len(files))(*[p.encode() for p in files])
parse = {}
for threads in (1, 8, 0):
    best = 1e9
    for _ in range(2):
        h, failed = C.c_void_p(), C.c_int(0)
        t0 = time.perf_counter()
        check(L.rv_replay_from_files(arr, len(files), 0, A.RULE_DEFAULT_TENHOU, threads, C.byref(h), C.byref(failed)))
        best = min(best, time.perf_counter() - t0)
        rounds = L.rv_replay_num_rounds(h)
        L.rv_replay_free(h)
    parse[str(threads or os.cpu_count())] = {"s": best, "files_per_sec": len(files) / best, "kyoku_per_sec": rounds / best}

warm = R.ReplayBatch.from_files(paths[:1], threads=1)  # warm-up: CUDA context, device tables, torch's lazily loaded kernels (once per process)
_o, _m, _i = (torch.empty((4 * warm.n, 74, 34), device="cuda"), torch.empty((4 * warm.n, 82), dtype=torch.uint8, device="cuda"),
              torch.empty((4 * warm.n,), dtype=torch.int32, device="cuda"))
for _ in range(3):
    _n = warm.vec.encode(obs=_o, mask=_m, index=_i)
    _ = (warm.labels_of_rows(_i, _n) >= 0).sum()
    warm.advance()
torch.cuda.synchronize()
t0 = time.perf_counter()
batch = R.ReplayBatch.from_files(files, threads=0)
torch.cuda.synchronize()
t_load = time.perf_counter() - t0
K = batch.n
# the same load with the action records staged in a pinned buffer the loader keeps across batches
pinned = torch.empty(int(batch._first[-1]) * C.sizeof(A.LogAction), dtype=torch.uint8).pin_memory()
del batch
torch.cuda.synchronize()
t0 = time.perf_counter()
batch = R.ReplayBatch.from_files(files, threads=0, staging=pinned)
torch.cuda.synchronize()
t_load_pinned = time.perf_counter() - t0
obs = torch.empty((2 * K, 74, 34), dtype=torch.float32, device="cuda")
mask = torch.empty((2 * K, 82), dtype=torch.uint8, device="cuda")
idx = torch.empty((2 * K,), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
rows = 0
labelled_dev = torch.zeros((), dtype=torch.int64, device="cuda")
while True:
    n = batch.vec.encode(obs=obs, mask=mask, index=idx)
    lab = batch.labels_of_rows(idx, n)
    rows += n
    labelled_dev += (lab >= 0).sum()              # (a consumer would gather obs[lab >= 0] here; no host sync per position)
    if not batch.advance():
        break
torch.cuda.synchronize()
t_walk = time.perf_counter() - t0
labelled = int(labelled_dev)
seat, aid = batch.labels()
assert labelled == int((aid >= 0).sum()), (labelled, int((aid >= 0).sum()))
print(json.dumps({
    "metric": "labelled_rows_per_sec", "value": labelled / (t_load_pinned + t_walk), "unit": "labelled decision rows/s (files -> tensors)", "n_gpus": 1,
    "config": {"workload": f"{len(files)} gzip MJAI logs ({G} distinct simulated 4p-red-half hanchan x {REP}), {K:,} kyoku: "
                           "ReplayBatch.from_files + advance / encode / labels_of_rows per position", "files": len(files), "kyoku": K,
               "host_threads": os.cpu_count()},
    "parse": parse, "load_s": t_load_pinned, "load_s_pageable": t_load, "value_pageable": labelled / (t_load + t_walk), "walk_s": t_walk, "rows": rows, "labelled_rows": labelled,
    "rows_per_sec_walk_only": labelled / t_walk,
}))
