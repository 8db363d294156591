#!/usr/bin/env python3
"""A/B of the headline rollout between builds / knob settings in ONE process on ONE box (fresh boxes differ by a few %):
usage: ab_rollout.py SPEC SPEC [SPEC ...] [--games N] — SPEC = lib.so[:VAR=val[,VAR=val...]] (environment knobs the library
reads per call, e.g. RV_ACT_HOLD=0); alternates the variants, 5 timed rollouts each, prints G env steps/s per rollout."""
import ctypes as C
import os
import sys

import torch  # noqa: F401  (brings the CUDA runtime libraries into the process)


def load(path):
    import os

    L = C.CDLL(os.path.abspath(path))
    vp = C.c_void_p
    L.rv_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.rv_vec_create.argtypes = [vp, C.c_int64, C.c_int, C.c_uint32, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint32, C.POINTER(vp)]
    L.rv_vec_reset.argtypes = [vp] + [vp] * 6
    L.rv_vec_reseed.argtypes = [vp, vp, C.c_uint64]
    L.rv_vec_step_random.argtypes = [vp, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]
    L.rv_timer_mark.argtypes = [vp, C.c_int]
    L.rv_timer_elapsed.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.rv_ctx_sync.argtypes = [vp]
    ctx, v = vp(), vp()
    assert L.rv_ctx_create(0, C.byref(ctx)) == 0
    return L, ctx


def rollout(L, ctx, v, k):
    L.rv_vec_reseed(v, None, 1000000 * k)
    L.rv_vec_reset(v, None, None, None, None, None, None)
    L.rv_ctx_sync(ctx)
    n = C.c_uint64(0)
    L.rv_timer_mark(ctx, 0)
    rc = L.rv_vec_step_random(v, 0x5EED, 1 << 30, C.byref(n))
    if rc != 0:
        L.rv_last_error.restype = C.c_char_p
        raise RuntimeError(f"rv_vec_step_random -> {rc}: {L.rv_last_error().decode()} (env: {[(k, v2) for k, v2 in os.environ.items() if k.startswith('RV_')]})")
    L.rv_timer_mark(ctx, 1)
    ms = C.c_float(0)
    L.rv_timer_elapsed(ctx, 0, 1, C.byref(ms))
    return n.value / ms.value / 1e6


import os

args = [a for a in sys.argv[1:] if not a.startswith("--")]
games = 65536
if "--games" in sys.argv:
    games = int(sys.argv[sys.argv.index("--games") + 1])
    args.remove(str(games))
libs, loaded = [], {}
for spec in args:
    path, _, knobs = spec.partition(":")
    env = dict(kv.split("=", 1) for kv in knobs.split(",") if kv)
    if path not in loaded:
        L, ctx = load(path)
        v = C.c_void_p()
        assert L.rv_vec_create(ctx, games, 2, 0xC0, None, 0, 0, C.byref(v)) == 0
        loaded[path] = (L, ctx, v)
    libs.append((spec, env) + loaded[path])
res = {spec: [] for spec, *_ in libs}
knob_names = sorted({k for _, env, *_ in libs for k in env})
for k in range(7):
    for spec, env, L, ctx, v in libs:
        for name in knob_names:
            os.environ.pop(name, None)
        os.environ.update(env)
        r = rollout(L, ctx, v, k)
        if k >= 2:
            res[spec].append(r)
for p, r in res.items():
    print(f"{p}: {' '.join(f'{x:.3f}' for x in r)}  median {sorted(r)[len(r) // 2]:.3f} G env steps/s")
