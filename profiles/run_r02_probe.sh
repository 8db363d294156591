#!/bin/bash
# Round 2, VERDICT item 2: probe the GPU box for a real riichienv build / Rust toolchain (wall golden vectors + Rust CPU baseline).
out=gpurun_out/r02_probe.txt
{
echo "== date"; date -u
echo "== import riichienv"; python -c "import riichienv; print(riichienv.__file__, getattr(riichienv,'__version__',None))" 2>&1 | tail -2
echo "== pip install riichienv==0.4.8 (no network expected)"; timeout 60 python -m pip install --target /tmp/rv_probe riichienv==0.4.8 2>&1 | tail -3
echo "== pip download"; timeout 60 python -m pip download -d /tmp/rv_dl riichienv==0.4.8 2>&1 | tail -3
echo "== wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "riichi|maturin|rust|pyo3" ; echo "(end)"
echo "== toolchains"; for t in cargo rustc rustup maturin go node javac clang; do printf "%s: " $t; command -v $t || echo MISSING; done
echo "== ~/.cargo"; ls -d /root/.cargo /usr/local/cargo /opt/rust* 2>&1 | tail -3
echo "== find riichienv artefacts"; find / \( -iname "*riichienv*" -o -iname "_riichienv*" \) -not -path "/proc/*" -not -path "$GRAFT_REPO_ROOT/*" -not -path "/root/repo/*" 2>/dev/null | head
echo "== baseline/_ref"; ls baseline/_ref 2>&1 | head
echo "== /root/reference"; ls /root/reference 2>&1 | head -3
echo "== cpu"; nproc; lscpu | grep -E "Model name|Socket|Thread|Core" 
echo "== gpu"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
} > $out 2>&1
cat $out
