#!/usr/bin/env python3
"""Headline benchmark: env steps/sec of seeded random-agent 4p hanchan (BASELINE.json).

One bench "step" = one pass of the hot path over one batch: G games per GPU are re-seeded,
reset (wall shuffle + deal) and played to `done` by the on-device keyed random agent.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA)
  python bench.py --impl reference --gpus N --steps K ...  # the reference arm: the CPU oracle
        (restatement of riichienv-core — the reference itself is Rust and cannot be built in
         this image) on all host cores, bounded sample per step

For N > 1 the driver launches this file under torchrun (one rank per GPU, NCCL).  Games are
independent: rank r owns a disjoint range of global game ids ("scaling": "weak"); the only
collective is the end-of-run reduction of timing / episode statistics.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_STEP = 1024  # algorithmic bytes per env step without observations (SURVEY.md §8 d)
MODE_NAMES = {0: "4p-red-single kyoku", 1: "4p-red-east", 2: "4p-red-half hanchan", 3: "3p-red-single kyoku", 4: "3p-red-east",
              5: "3p-red-half hanchan (sanma)"}
METRIC = "env_steps_per_sec"
UNIT = "env steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=65536, help="games per GPU per bench step")
    ap.add_argument("--mode", type=int, default=2, help="2 = 4p-red-half")
    ap.add_argument("--cpu-sample-games", type=int, default=0, help="games in the cpu_baseline sample (0 = sized by time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--split", action="store_true", help="rollout_obs: tensor rows by rv_vec_encode, masks + step by the fused kernel")
    ap.add_argument("--unfused", action="store_true", help="rollout_obs: rv_vec_encode + rv_vec_step_random instead of the fused kernel")
    ap.add_argument("--workload", default="rollout", choices=["rollout", "rollout_obs", "hands"],
                    help="rollout: BASELINE configs[2]/[3] (headline); rollout_obs: configs[4] (encode()+mask() every step); "
                         "hands: configs[1] (batched shanten + agari/yaku/fu/score)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def source_fingerprint():
    """sha1 over the CUDA sources (kernels and the headers they include; the host-only .cpp files — JSON renderer, log reader —
    cannot change a kernel): ncu facts are only quoted for the build they were measured on"""
    import hashlib

    h = hashlib.sha1()
    d = os.path.join(ROOT, "riichienv_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_facts(kernel):
    """DRAM bytes per launch and issue-slot utilisation of `kernel` from the committed `ncu --set full` capture of this build
    (profiles/ncu_facts.json, written by profiles/summarize_ncu.py --facts); None when the capture belongs to other sources,
    so the numbers cannot go stale silently."""
    try:
        facts = json.load(open(os.path.join(ROOT, "profiles", "ncu_facts.json")))
    except Exception:
        return None
    f = facts.get(kernel)
    if not f or f.get("src") != source_fingerprint():
        return None
    return f


def workload_config(mode, games_per_gpu):
    return {"workload": (f"{MODE_NAMES[mode]}, default Tenhou rules, {games_per_gpu:,} parallel seeded random-agent games per GPU "
                         f"(BASELINE.json configs[{3 if mode >= 3 else 2}]), reset -> done"),
            "games_per_gpu": games_per_gpu, "game_mode": mode}


def cpu_sample_games(args):
    """hanchan per CPU step: the same bounded sample for the cpu_baseline leg and the --impl reference arm"""
    return args.cpu_sample_games or max(64 * (os.cpu_count() or 1), 2048)


class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 - 0.2 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_oracle_sample(mode, n_games, seed_base, agent_seed, threads):
    import numpy as np

    import oracle
    from riichienv_b200 import _abi as A

    lib = oracle.load()
    t0 = time.perf_counter()
    steps = lib.orc_run_random(mode, A.RULE_DEFAULT_TENHOU, seed_base, n_games, agent_seed, 1 << 30, threads, None, None, None, None,
                               None, None, None, None)
    return int(steps), time.perf_counter() - t0


def cpu_baseline(mode, sample_games, reps=6):
    """the oracle port on every host core: `reps` steps of the same bounded sample the --impl reference arm plays per step"""
    threads = os.cpu_count() or 1
    run_oracle_sample(mode, max(threads, sample_games // 8), 10_000_000, 1, threads)      # warm-up
    steps = 0
    dt = 0.0
    for k in range(reps):
        s_, d_ = run_oracle_sample(mode, sample_games, 20_000_000 + k * sample_games, 1, threads)
        steps += s_
        dt += d_
    return {"value": steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} x {sample_games} seeded {MODE_NAMES[mode]} ({steps} env steps, {dt:.1f} s) through the C++ oracle "
                      f"port (oracle/: a restatement of riichienv-core; the Rust reference cannot be built here or on the GPU box, "
                      f"profiles/r02_probe.txt), {threads} threads"}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = cpu_sample_games(args)                              # ~1-2 s of work per step on all cores
    for w in range(args.warmup):
        run_oracle_sample(args.mode, max(threads, per_step // 8), 30_000_000 + w * per_step, 1, threads)
    tot_steps, tot_t = 0, 0.0
    for k in range(args.steps):
        s, dt = run_oracle_sample(args.mode, per_step, 40_000_000 + k * per_step, 1, threads)
        tot_steps += s
        tot_t += dt
    val = tot_steps / tot_t
    sample = (f"{per_step} seeded hanchan per step x {args.steps} steps ({tot_steps} env steps) through the C++ oracle port "
              f"(CPU restatement of riichienv-core; no Rust toolchain here or on the GPU box, so oracle/_ref cannot exist), "
              f"{threads} threads")
    cfg = workload_config(args.mode, args.games)
    cfg.update({"cpu_arm": "oracle port (kind=port), not riichienv-core itself", "sample_games_per_step": per_step})
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/i32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch

    from riichienv_b200._lib import Context
    from riichienv_b200.multi_gpu import RunStats, reduce_stats, shard_range
    from riichienv_b200.vec_env import VecRiichiEnv

    if args.workload != "rollout":
        import bench_extra

        return bench_extra.run(args, rank, world, local_rank)

    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    torch.cuda.set_device(local_rank)
    ctx = Context.get(local_rank)
    ext_stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    G = args.games
    v = VecRiichiEnv(G, args.mode, device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2
    agent_seed = 0x5EED

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def one_step(k, timed):
        """returns (step_ms, kernel_ms, env_steps)"""
        base, _ = shard_range(k, world, rank, G)  # disjoint global game ids per (bench step, rank)
        with torch.cuda.stream(ext_stream):
            flush.fill_(k & 0xFF)              # L2 flush between iterations (outside the timed events)
        s_before, _ = v.steps_total()
        ctx.timer_mark(0)
        v.reseed(None, base)
        v.reset()                              # synchronises the stream internally
        ctx.timer_mark(1)
        v.step_random_async(agent_seed, 1 << 30)
        ctx.timer_mark(2)
        kernel_ms = ctx.timer_elapsed(1, 2)
        step_ms = ctx.timer_elapsed(0, 2)
        s_after, _ = v.steps_total()
        # reset zeroes the device counter, so s_after is this step's count
        return step_ms, kernel_ms, s_after

    for w in range(args.warmup):
        one_step(1000 + w, False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    t0 = time.time()
    tot_ms = tot_kernel_ms = 0.0
    tot_steps = 0
    launches = 0
    for k in range(args.steps):
        ms, kms, st = one_step(k, True)
        tot_ms += ms
        tot_kernel_ms += kms
        tot_steps += st
        launches += 4                         # reseed_kernel + reset_kernel + q_init_kernel + rollout_crew_kernel
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1)
    done, scores, ranks = v.results()
    assert done.all(), "a game did not finish inside the timed region"

    # ---- e2e: the public host-buffer API, H2D seeds in and D2H results out every step ----------
    pinned = torch.empty(G, dtype=torch.int64).pin_memory()
    e2e_t = 0.0
    e2e_steps = 0
    for k in range(args.steps):
        base, _ = shard_range(5000 + k, world, rank, G)
        pinned.copy_(torch.arange(base, base + G, dtype=torch.int64))
        seeds = pinned.numpy().view(np.uint64)
        ctx.sync()
        a = time.perf_counter()
        v.reseed(seeds, 0)                     # H2D: 8 B / game
        v.reset()
        t_r = time.perf_counter()
        n = v.step_random(agent_seed, 1 << 30)
        t_s = time.perf_counter()
        d_, s_, r_ = v.results()               # D2H: done + scores + ranks
        c_ = v.counters()                      # D2H: step / kyoku / event counters + event hash
        e2e_t += time.perf_counter() - a
        if os.environ.get("RV_DEBUG"):
            print(f"[e2e] reseed+reset {1e3 * (t_r - a):.1f} ms, rollout {1e3 * (t_s - t_r):.1f} ms, results+counters "
                  f"{1e3 * (time.perf_counter() - t_s):.1f} ms", file=sys.stderr)
        e2e_steps += n
    h2d = G * 8
    d2h = G * (1 + 16 + 4 + 4 + 4 + 4 + 8)

    # ---- reduce over ranks: max time, sum of steps (the only collective) -----------------------
    mine = RunStats(elapsed_ms=tot_ms, kernel_ms=tot_kernel_ms, e2e_s=e2e_t, env_steps=float(tot_steps),
                    e2e_steps=float(e2e_steps), games=float(G * args.steps), score_sum=float(scores.sum()))
    red = reduce_stats(mine, dist, torch, f"cuda:{local_rank}")
    tot_ms, e2e_t = red.elapsed_ms, red.e2e_s
    all_steps, all_e2e_steps = red.env_steps, red.e2e_steps

    if rank == 0:
        peak, peak_src = measured_peaks()
        value = all_steps / (tot_ms / 1000.0)
        achieved = (tot_steps * B_STEP) / (tot_kernel_ms / 1000.0) / 1e9  # this rank's dominant kernel
        kname = "rollout_crew_kernel<3>" if args.mode >= 3 else "rollout_crew_kernel<4>"
        facts = ncu_facts(kname) if (G == 65536 and args.mode in (2, 5)) else None
        cfg = workload_config(args.mode, G)
        cfg.update({"l2": "256 MiB flush write between timed iterations",
                    "games_per_sec": (G * args.steps * world) / (tot_ms / 1000.0),
                    "env_steps_per_game": all_steps / (G * args.steps * world)})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_ms / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/i32", "data": "synthetic",
            "config": cfg,
            "e2e": {"value": all_e2e_steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # traffic / issue_frac: from the committed ncu --set full capture of THIS build (else null)
                         "traffic": facts["dram_bytes"] if facts else None,
                         "issue_frac": facts["issue_active_pct"] / 100.0 if facts else None,
                         "ncu_capture": facts["capture"] if facts else None,
                         "kernel": "rollout_crew_kernel (one launch per rollout)", "peak_source": peak_src,
                         "bytes_per_env_step": B_STEP, "kernel_share_of_step": tot_kernel_ms / tot_ms},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.mode, cpu_sample_games(args))
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
