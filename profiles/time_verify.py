#!/usr/bin/env python3
"""MjSoulReplay.verify measured: every win of G paifu games checked against what the paifu recorded.

Input: hanchan played by this repo's simulator (oracle, greedy-win agent), written as paifu rounds.  Timed:
  read      rv_replay_from_mjsoul_text over the games' JSON (host)
  walk      rv_replay_win_contexts over every round (host: the WinResultContextIterator walk -> one rv_hand_query per win)
  evaluate  ONE rv_hand_eval_batch over the queries of all games, host buffers (CUDA events around the call)
and the oracle's evaluator over the same queries on one host thread.  The query list is tiled to N for the batch figure.
usage: time_verify.py [games] [N]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import oracle  # noqa: E402
from riichienv_b200 import _abi as A  # noqa: E402
from riichienv_b200._lib import Context, check, lib  # noqa: E402
import riichienv_b200.replay as R  # noqa: E402
from tests.test_replay import _paifu_rounds, simulated_log  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000

texts = []
for seed in range(G):
    rounds = [ev for _, ev in _paifu_rounds(R, "\n".join(simulated_log(2, 700 + seed)) + "\n", False)]
    texts.append(json.dumps({"rounds": rounds}).encode())

L = lib()
t0 = time.perf_counter()
handles = []
for t in texts:
    h = C.c_void_p()
    check(L.rv_replay_from_mjsoul_text(t, len(t), A.RULE_DEFAULT_MJSOUL, C.byref(h)))
    handles.append(h)
t_read = time.perf_counter() - t0
t0 = time.perf_counter()
ctxs = []
n_rounds = 0
for h in handles:
    n = C.c_int(0)
    check(L.rv_replay_win_contexts(h, -1, None, 0, C.byref(n)))
    arr = (A.WinContext * max(1, n.value))()
    check(L.rv_replay_win_contexts(h, -1, arr, n.value, C.byref(n)))
    ctxs += [arr[i] for i in range(n.value)]
    n_rounds += L.rv_replay_num_rounds(h)
t_walk = time.perf_counter() - t0
for h in handles:
    L.rv_replay_free(h)
W = len(ctxs)
reps = (N + W - 1) // W
base = (A.HandQuery * W)(*[c.query for c in ctxs])
q = (A.HandQuery * (W * reps)).from_buffer_copy(bytes(base) * reps)
n = W * reps
out = (A.HandResult * n)()
ctx = Context.get(0)
check(L.rv_hand_eval_batch(ctx.handle, q, out, n))          # warm-up
best = 1e9
for _ in range(3):
    ctx.sync()
    ctx.timer_mark(0)
    check(L.rv_hand_eval_batch(ctx.handle, q, out, n))
    ctx.timer_mark(1)
    ctx.sync()
    best = min(best, ctx.timer_elapsed(0, 1))
want = (A.HandResult * W)()
t0 = time.perf_counter()
oracle.load().orc_hand_eval(base, want, W)
t_orc = time.perf_counter() - t0
assert bytes(out)[: C.sizeof(A.HandResult) * W] == bytes(want), "device result differs from the oracle"
wins = sum(1 for i in range(W) if out[i].is_win)
print(json.dumps({
    "metric": "wins_verified_per_sec", "value": n / (best / 1e3), "unit": "wins/s", "n_gpus": 1,
    "config": {"workload": f"{G} simulated 4p-red-half hanchan as paifu ({n_rounds} rounds, {W} wins), queries tiled to {n:,}; "
                           "one rv_hand_eval_batch with host buffers", "wins": W, "batch": n},
    "ms_batch": best, "is_win": wins,
    "host": {"read_ms": t_read * 1e3, "walk_ms": t_walk * 1e3, "rounds_per_sec_walk": n_rounds / t_walk,
             "note": "rv_replay_from_mjsoul_text and rv_replay_win_contexts on one host thread (ctypes call per game)"},
    "cpu_baseline": {"value": W / t_orc, "unit": "wins/s", "cores": 1, "kind": "port",
                     "sample": f"the {W} distinct queries through the oracle's evaluator"},
}))
