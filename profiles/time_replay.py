#!/usr/bin/env python3
"""Replay ingestion measured: K kyoku replayed in lock-step on one vector of K records.

Input: MJAI logs of hanchan played by this repo's own simulator (oracle, greedy-win agent), parsed by the library's reader
(rv_replay_from_text) and tiled to K kyoku.  Timed, with CUDA events around the whole replay (host buffers, H2D of every action
array inside the timed region — the C-ABI call a data loader makes):
  tracking      rv_vec_replay_begin + one rv_vec_apply_log_actions per log position
  + tensors     the same with rv_vec_encode (74 x 34 f32 rows + 82-id masks of every seat that owes a decision) after every position
and the oracle's apply_log_action over the same kyoku on one host thread (a bounded sample).  Prints one JSON line.
usage: time_replay.py [K] [hanchan]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from riichienv_b200 import _abi as A  # noqa: E402
from riichienv_b200._lib import Context, check, lib  # noqa: E402
from riichienv_b200.vec_env import VecRiichiEnv  # noqa: E402
from tests.test_replay import OracleBackend, begin, parse_text, simulated_log  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
GAMES = int(sys.argv[2]) if len(sys.argv) > 2 else 16

rounds = []
for seed in range(GAMES):
    rounds += parse_text("\n".join(simulated_log(2, 900 + seed)) + "\n")
nk = len(rounds)
reps = (K + nk - 1) // nk
K = reps * nk
T = max(len(a) for _, a in rounds)
n_actions = sum(len(a) for _, a in rounds) * reps
ky = (A.LogKyoku * K).from_buffer_copy(bytes((A.LogKyoku * nk)(*[k for k, _ in rounds])) * reps)
steps = []
for t in range(T):
    base = (A.LogAction * nk)()
    for i, (_, acts) in enumerate(rounds):
        if t < len(acts):
            base[i] = acts[t]
    steps.append((A.LogAction * K).from_buffer_copy(bytes(base) * reps))

ctx = Context.get(0)
v = VecRiichiEnv(K, 0, A.RULE_DEFAULT_TENHOU, seed_base=0, log_cap_words=0)
obs = torch.empty((2 * K, 74, 34), dtype=torch.float32, device="cuda")
mask = torch.empty((2 * K, 82), dtype=torch.uint8, device="cuda")
idx = torch.empty((2 * K,), dtype=torch.int32, device="cuda")


first = (C.c_int64 * (K + 1))()
lens = [len(a) for _, a in rounds]
for i in range(K):
    first[i + 1] = first[i] + lens[i % nk]
flat = (A.LogAction * sum(lens))(*[a for _, acts in rounds for a in acts])
all_actions = (A.LogAction * first[K]).from_buffer_copy(bytes(flat) * reps)


# the concatenated logs once more in PINNED host memory (a data loader's staging buffer)
_pin = torch.frombuffer(bytearray(bytes(all_actions)), dtype=torch.uint8).pin_memory()
pinned_actions = C.cast(_pin.data_ptr(), C.POINTER(A.LogAction))
resident_src = {"pinned": False}


def replay_resident(with_rows):
    """the log uploaded once (rv_vec_replay_load), then one kernel launch per position"""
    rows = 0
    v.replay_load(ky, pinned_actions if resident_src["pinned"] else all_actions, first)
    while v.replay_advance() != 0:
        if with_rows:
            rows += v.encode(obs=obs, mask=mask, index=idx)
    return rows


def replay(with_rows):
    rows = 0
    v.replay_begin(ky)
    for arr in steps:
        v.apply_log_actions(arr)
        if with_rows:
            rows += v.encode(obs=obs, mask=mask, index=idx)
    return rows


def timed(with_rows, n=3, replay=replay):
    replay(with_rows)                       # warm-up
    best, rows = 1e9, 0
    for _ in range(n):
        ctx.sync()
        ctx.timer_mark(0)
        rows = replay(with_rows)
        ctx.timer_mark(1)
        ctx.sync()
        best = min(best, ctx.timer_elapsed(0, 1))
    return best, rows


ms_track, _ = timed(False)
ms_rows, rows = timed(True)
ms_res, _ = timed(False, replay=replay_resident)
ms_res_rows, rows_res = timed(True, replay=replay_resident)
assert rows_res == rows
resident_src["pinned"] = True
ms_pin, _ = timed(False, replay=replay_resident)
ms_pin_rows, _ = timed(True, replay=replay_resident)

# the oracle on one host thread, a bounded sample of the distinct kyoku
t0 = time.perf_counter()
o_actions = 0
for k, acts in rounds[: min(nk, 64)]:
    b = begin(OracleBackend, k)
    for a in acts:
        b.apply(a)
    o_actions += len(acts)
o_s = time.perf_counter() - t0

print(json.dumps({
    "metric": "log_actions_per_sec", "value": n_actions / (ms_track / 1e3), "unit": "log actions/s", "n_gpus": 1,
    "config": {"workload": f"{K:,} kyoku ({nk} distinct, from {GAMES} simulated 4p-red-half hanchan) replayed in lock-step: "
                           f"rv_vec_replay_begin + {T} x rv_vec_apply_log_actions with host buffers", "kyoku": K, "positions": T,
               "log_actions": n_actions},
    "ms_tracking": ms_track,
    "resident": {"log_actions_per_sec": n_actions / (ms_res / 1e3), "ms_tracking": ms_res, "ms_with_tensors": ms_res_rows,
                 "rows_per_sec": rows / (ms_res_rows / 1e3),
                 "note": "rv_vec_replay_load (ONE upload of all logs, inside the timed region) + one rv_vec_replay_advance per position",
                 "pinned": {"log_actions_per_sec": n_actions / (ms_pin / 1e3), "ms_tracking": ms_pin, "ms_with_tensors": ms_pin_rows,
                            "rows_per_sec": rows / (ms_pin_rows / 1e3), "note": "the same with the logs in pinned host memory"}},
    "with_tensors": {"ms": ms_rows, "rows": rows, "rows_per_sec": rows / (ms_rows / 1e3),
                     "note": "rv_vec_encode after every position: 74x34 f32 + 82-id mask per seat that owes a decision"},
    "h2d_bytes": K * (C.sizeof(A.LogKyoku) + T * C.sizeof(A.LogAction)),
    "cpu_baseline": {"value": o_actions / o_s, "unit": "log actions/s", "cores": 1, "kind": "port",
                     "sample": f"{o_actions} actions of {min(nk, 64)} kyoku through the oracle's apply_log_action (ctypes call per action)"},
}))
