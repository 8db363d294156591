#!/bin/bash
# round 2: compute-sanitizer racecheck + synccheck + memcheck on the scheduler kernels at small n
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from riichienv_b200.vec_env import VecRiichiEnv
which = sys.argv[1]
n = int(sys.argv[2])
v = VecRiichiEnv(n, 2, seed_base=11)
v.reset()
if which == "persist":
    t = v.step_random(5, 400)          # persistent class-queue rollout (max_steps >= 32)
elif which == "lockstep":
    t = sum(v.step_random(5, 1) for _ in range(60))     # step_staged_kernel
else:
    obs = torch.empty((2 * n, 74, 34), dtype=torch.float32, device="cuda")
    mask = torch.empty((2 * n, 82), dtype=torch.uint8, device="cuda")
    idx = torch.empty((2 * n,), dtype=torch.int32, device="cuda")
    for _ in range(40):
        v.observe_step_random(5, obs=obs, mask=mask, index=idx, sync=True)
    t = v.steps_total()[0]
print(which, n, "steps", t)
PY
for tool in racecheck synccheck memcheck; do
for w in persist lockstep observe; do
  n=512; [ $w = persist ] && n=4096
  echo "== $tool $w n=$n" >> gpurun_out/r02_sanitize.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_small.py $w $n 2>&1 | grep -v "^$" | tail -6 >> gpurun_out/r02_sanitize.txt
done
done
cat gpurun_out/r02_sanitize.txt
