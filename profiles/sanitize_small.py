#!/usr/bin/env python3
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): rollouts (4P, sanma), lock-step steps,
encode / encode_extended / encode_seq / observe+step, hand evaluation.  Sizes are tiny: the tools slow kernels 10-100x."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from riichienv_b200 import _abi as A
from riichienv_b200._lib import Context, check, lib
from riichienv_b200.vec_env import VecRiichiEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for mode in (2, 5):
    v = VecRiichiEnv(n, mode, A.RULE_DEFAULT_TENHOU, seed_base=100, log_cap_words=1 << 14)
    v.reset()
    W, IDS = (27, 60) if mode >= 3 else (34, 82)
    obs = torch.empty((n * 3, 74, W), dtype=torch.float32, device="cuda")
    mask = torch.empty((n * 3, IDS), dtype=torch.uint8, device="cuda")
    idx = torch.empty((n * 3,), dtype=torch.int32, device="cuda")
    for it in range(40):
        v.encode(obs=obs, mask=mask, index=idx)
        v.observe_step_random(3, obs=obs, mask=mask, index=idx, sync=True)
        v.step_random(3, 7)
    if mode == 2:
        ext = torch.empty((n * 3, 215, 34), dtype=torch.float32, device="cuda")
        v.encode_extended(obs=ext, mask=mask, index=idx)
        sp = torch.zeros((n * 3, 25), dtype=torch.uint16, device="cuda")
        nu = torch.zeros((n * 3, 12), dtype=torch.float32, device="cuda")
        pr = torch.zeros((n * 3, 64, 5), dtype=torch.uint16, device="cuda")
        ca = torch.zeros((n * 3, 64, 4), dtype=torch.uint16, device="cuda")
        le = torch.zeros((n * 3, 3), dtype=torch.uint16, device="cuda")
        v.encode_seq(sparse=sp, numeric=nu, prog=pr, cand=ca, lens=le, index=idx)
    total = v.step_random(3, 100000)
    done, scores, ranks = v.results()
    assert done.all(), mode
    print("mode", mode, "steps", total, "score sum", int(scores.sum()))
from tests import helpers as H

cases = H.load_agari_cases()[:256]
arr = H.query_array([c[0] for c in cases])
out = (A.HandResult * len(cases))()
check(lib().rv_hand_eval_batch(Context.get(0).handle, arr, out, len(cases)))
print("sanitize_small ok")
