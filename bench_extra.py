"""Secondary benchmark workloads (same JSON contract as bench.py; not the headline):

  --workload rollout_obs   BASELINE.json configs[4]: 4p-red-half rollout with FEATURE_ENCODING tensors
                           (74x34 f32) + 82-id masks written for every acting seat at every env step
  --workload hands         BASELINE.json configs[1]: batched shanten + agari/yaku/fu/score over seeded hands
"""
import ctypes as C
import json
import os
import time

B_OBS = 74 * 34 * 4 + 82   # algorithmic bytes per observation written (SURVEY.md §8 d)
B_HAND = 56 + 40           # rv_hand_query in + rv_hand_result out


def _peak():
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def run(args, rank, world, local_rank):
    import numpy as np
    import torch

    from riichienv_b200 import _abi as A
    from riichienv_b200._lib import Context, check, lib
    from riichienv_b200.multi_gpu import RunStats, reduce_stats, shard_range
    from riichienv_b200.vec_env import VecRiichiEnv

    dist = None
    if world > 1:
        assert args.workload == "rollout_obs", "the hands workload is single-GPU"
        import torch.distributed as dist_mod

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    torch.cuda.set_device(local_rank)
    ctx = Context.get(local_rank)
    peak, peak_src = _peak()
    if args.workload == "hands":
        n = 10_000_000
        # 10^7 DISTINCT seeded hands (include/rv_synth.h: hand h from splitmix64(0xA6A21 + h); every tenth hand is a synthetic
        # complete hand), generated on the device
        d_q = torch.empty((n, C.sizeof(A.HandQuery)), dtype=torch.uint8, device="cuda")
        d_r = torch.empty((n, C.sizeof(A.HandResult)), dtype=torch.uint8, device="cuda")
        check(lib().rv_hand_queries_seeded(ctx.handle, C.c_void_p(d_q.data_ptr()), 0, n))
        ctx.sync()
        base = list(range(100_000))
        arr = d_q[: len(base)].cpu().numpy()
        ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

        def go():
            check(lib().rv_hand_eval_batch_device(ctx.handle, C.c_void_p(d_q.data_ptr()), C.c_void_p(d_r.data_ptr()), n))

        for _ in range(args.warmup):
            go()
        ctx.sync()
        ctx.timer_mark(0)
        for _ in range(args.steps):
            go()
        ctx.timer_mark(1)
        ms = ctx.timer_elapsed(0, 1)
        # e2e: host buffers through rv_hand_eval_batch (H2D + kernel + D2H inside the call)
        ne = 2_000_000                                           # hands per e2e call, ordinary (pageable) host memory
        h_in = np.ascontiguousarray(np.tile(arr, (ne // len(base), 1)))
        h_out = np.empty((ne, C.sizeof(A.HandResult)), np.uint8)
        pq, po = h_in.ctypes.data_as(C.POINTER(A.HandQuery)), h_out.ctypes.data_as(C.POINTER(A.HandResult))
        check(lib().rv_hand_eval_batch(ctx.handle, pq, po, ne))   # warm-up: allocates the staging buffers
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            check(lib().rv_hand_eval_batch(ctx.handle, pq, po, ne))
        e2e_pageable = reps * ne / (time.perf_counter() - t0)
        # the same call on PINNED host buffers (what the bench contract asks for; the library sees that the caller's memory is
        # page-locked and copies straight from / into it)
        p_in = torch.from_numpy(h_in).pin_memory()
        p_out = torch.empty((ne, C.sizeof(A.HandResult)), dtype=torch.uint8).pin_memory()
        ppq, ppo = C.cast(p_in.data_ptr(), C.POINTER(A.HandQuery)), C.cast(p_out.data_ptr(), C.POINTER(A.HandResult))
        check(lib().rv_hand_eval_batch(ctx.handle, ppq, ppo, ne))
        assert bytes(p_out.numpy()[:4096].tobytes()) == h_out[:4096].tobytes()
        t0 = time.perf_counter()
        for _ in range(reps):
            check(lib().rv_hand_eval_batch(ctx.handle, ppq, ppo, ne))
        e2e = reps * ne / (time.perf_counter() - t0)
        val = n * args.steps / (ms / 1000)
        ach = val * B_HAND / 1e9
        print(json.dumps({
            "metric": "hands_per_sec", "value": val, "unit": "hands/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32",
            "data": "synthetic",
            "config": {"workload": "10^7 distinct seeded 14-tile hands (include/rv_synth.h: splitmix64(0xA6A21 + h); 90 % uniform "
                                   "draws, 10 % synthetic complete hands): shanten(14), shanten(13), waits, agari, yaku/han/fu/"
                                   "score (BASELINE.json configs[1])", "hands": n,
                       "l2": "inputs+outputs 960 MB per step > 126 MB L2"},
            "e2e": {"value": e2e, "unit": "hands/s", "h2d_bytes_per_step": ne * 56, "d2h_bytes_per_step": ne * 40,
                    "note": "rv_hand_eval_batch on pinned host buffers, 2M hands per call", "pageable": e2e_pageable},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "kernel": "hand_eval_kernel", "peak_source": peak_src, "bytes_per_hand": B_HAND},
        }))
        return
    # ---- rollout with observations every step
    G = args.games
    v = VecRiichiEnv(G, args.mode, device=local_rank)
    max_obs = G * 2
    W, IDS = (27, 60) if args.mode >= 3 else (34, 82)     # sanma: Observation3P (74 x 27, 60 ids)
    b_obs = 74 * W * 4 + IDS
    obs = torch.empty((max_obs, 74, W), dtype=torch.float32, device="cuda")
    mask = torch.empty((max_obs, IDS), dtype=torch.uint8, device="cuda")
    idx = torch.empty((max_obs,), dtype=torch.int32, device="cuda")

    def one(k):
        v.reseed(None, shard_range(k, world, rank, G)[0])   # disjoint global game ids per (bench step, rank)
        v.reset()
        n_obs = 0
        it = 0
        while True:
            sync = (it % 64) == 63
            if args.unfused:
                r = v.encode(obs=obs, mask=mask, index=idx, max_obs=max_obs, sync=sync)
                v.step_random_async(0x5EED, 1)
            elif args.split:
                r = v.encode(obs=obs, index=idx, max_obs=max_obs, sync=sync)
                v.observe_step_random(0x5EED, mask=mask, max_obs=max_obs)
            else:
                r = v.observe_step_random(0x5EED, obs=obs, mask=mask, index=idx, max_obs=max_obs, sync=sync)
            it += 1
            if sync:
                n_obs += r * 64          # sampled row count (rows/iteration changes slowly)
                if r == 0:
                    break
        return it, n_obs

    for w in range(max(1, args.warmup // 3)):
        one(1000 + w)
    ctx.sync()
    if dist is not None:
        dist.barrier()
    tot_ms, tot_steps, tot_obs, iters, wall_s = 0.0, 0, 0, 0, 0.0
    for k in range(args.steps):
        ctx.sync()
        t_host = time.perf_counter()
        ctx.timer_mark(0)
        it, n_obs = one(k)
        ctx.timer_mark(1)
        tot_ms += ctx.timer_elapsed(0, 1)
        wall_s += time.perf_counter() - t_host        # e2e: the same public calls on the host clock (launches, syncs, row counts)
        s, _ = v.steps_total()
        tot_steps += s
        tot_obs += n_obs
        iters += it
    # the dominant kernel alone: tensor rows of a mid-game decision point through rv_vec_encode (count + scan + obs_encode_kernel),
    # CUDA events on the context stream, rows >> L2
    v.reseed(None, shard_range(9000, world, rank, G)[0])
    v.reset()
    v.step_random_async(0x5EED, 300)
    ctx.sync()
    reps = 20
    rows = v.encode(obs=obs, index=idx, max_obs=max_obs, sync=True)
    ctx.timer_mark(2)
    for _ in range(reps):
        v.encode(obs=obs, index=idx, max_obs=max_obs, sync=False)
    ctx.timer_mark(3)
    enc_ms = ctx.timer_elapsed(2, 3) / reps
    enc_gbs = rows * (74 * W * 4) / (enc_ms / 1000) / 1e9
    ach = (tot_steps * 1024 + tot_obs * b_obs) / (tot_ms / 1000) / 1e9      # this rank's whole pipeline
    red = reduce_stats(RunStats(elapsed_ms=tot_ms, env_steps=float(tot_steps), games=float(tot_obs), e2e_s=wall_s, e2e_steps=float(tot_steps)),
                       dist, torch, f"cuda:{local_rank}")
    tot_ms, all_steps, all_obs = red.elapsed_ms, red.env_steps, red.games      # max time over ranks, summed steps / rows
    wall_s = red.e2e_s
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return
    val = all_steps / (tot_ms / 1000)
    print(json.dumps({
        "metric": "env_steps_per_sec", "value": val, "unit": "env steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32 + f32 obs",
        "data": "synthetic",
        "config": {"workload": f"{'3p-red-half (sanma)' if args.mode >= 3 else '4p-red-half'} hanchan, {G:,} games per GPU, encode() (74x{W} f32) + mask() for every acting seat at every "
                               "env step (BASELINE.json configs[4]; games sharded over the GPUs)", "games_per_gpu": G,
                   "observations_per_env_step": all_obs / max(1, all_steps), "l2": "observation buffer 1.3 GB per iteration > L2"},
        "e2e": {"value": all_steps / wall_s, "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * (iters // 64) // max(1, args.steps),
                "note": "the same rv_vec_observe_step_random loop on the host clock (ctypes launches, a row-count read-back every 64 "
                        "iterations); the tensors are consumed on the device, as by the reference's GPU policy workers"},
        "gpu_launches": iters * (8 if not args.unfused else 6),
        "roofline": {"bound": "hbm", "achieved": enc_gbs, "peak": peak, "unit": "GB/s", "frac": enc_gbs / peak, "traffic": None,
                     "kernel": "obs_encode_kernel (dominant: tensor rows, timed alone through rv_vec_encode)", "peak_source": peak_src,
                     "bytes_per_observation": b_obs, "rows_per_launch": rows, "kernel_ms": enc_ms,
                     "kernel_share_of_step": enc_ms * iters / max(1e-9, tot_ms),
                     "pipeline_achieved": ach, "pipeline_frac": ach / peak},
    }))
