#!/usr/bin/env python3
"""Times rv_vec_encode_ext (215x34 rows) and rv_vec_encode (74x34 rows) on 65,536 mid-game hanchan (CUDA events, 10 reps).
usage: time_encode_ext.py [games] [mode]   (mode 5 = sanma: 215x27 / 74x27 rows, 60 mask ids)"""
import json
import sys

import torch

sys.path.insert(0, ".")
from riichienv_b200 import _abi as A
from riichienv_b200.vec_env import VecRiichiEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, IDS = (27, 60) if mode >= 3 else (34, 82)
v = VecRiichiEnv(n, mode, A.RULE_DEFAULT_TENHOU, seed_base=0)
v.reset()
v.step_random(1, 200 if mode >= 3 else 300)
cap = n + n // 4
ext = torch.empty((cap, 215, W), dtype=torch.float32, device="cuda")
base = torch.empty((cap, 74, W), dtype=torch.float32, device="cuda")
mask = torch.empty((cap, IDS), dtype=torch.uint8, device="cuda")
idx = torch.empty((cap,), dtype=torch.int32, device="cuda")
out = {}
for name, fn, buf in (("encode_ext", v.encode_extended, ext), ("encode", v.encode, base)):
    rows = fn(obs=buf, mask=mask, index=idx)
    ts = []
    for _ in range(10):
        v.ctx.timer_mark(0)          # CUDA events on the library's own stream
        fn(obs=buf, mask=mask, index=idx, sync=False)
        v.ctx.timer_mark(1)
        ts.append(v.ctx.timer_elapsed(0, 1))
    ms = sorted(ts)[len(ts) // 2]
    out[name] = {"rows": rows, "ms": ms, "rows_per_s": rows / ms * 1e3, "GBps": rows * buf[0].numel() * 4 / ms / 1e6}
out["mode"] = mode
print(json.dumps(out))
