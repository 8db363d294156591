"""Times the phases of the end-to-end call sequence bench.py uses for `e2e` (host buffers in, host results out)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from riichienv_b200.vec_env import VecRiichiEnv
from riichienv_b200 import _abi as A
G = 65536
v = VecRiichiEnv(G, 2, A.RULE_DEFAULT_TENHOU, seed_base=0)
ctx = v.ctx
for it in range(4):
    seeds = np.arange(it * G, (it + 1) * G, dtype=np.uint64)
    ctx.sync()
    t = [time.perf_counter()]
    v.reseed(seeds, 0); ctx.sync(); t.append(time.perf_counter())
    v.reset(); ctx.sync(); t.append(time.perf_counter())
    n = v.step_random(0x5EED, 1 << 30); t.append(time.perf_counter())
    v.results(); t.append(time.perf_counter())
    v.counters(); t.append(time.perf_counter())
    names = ["reseed", "reset", "step_random", "results", "counters"]
    print(it, n, " ".join(f"{nm}={1e3 * (b - a):.1f}ms" for nm, a, b in zip(names, t, t[1:])))
