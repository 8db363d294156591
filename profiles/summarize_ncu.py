#!/usr/bin/env python3
"""Key counters of one `ncu --set full` capture, from `ncu -i X.ncu-rep --page raw --csv` output.
usage: summarize_ncu.py raw.csv [row]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
vals = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
col = {h: i for i, h in enumerate(hdr)}
KEYS = """Kernel Name
gpu__time_duration.sum
launch__grid_size
launch__block_size
launch__registers_per_thread
launch__shared_mem_per_block_static
launch__shared_mem_per_block_dynamic
launch__occupancy_limit_shared_mem
launch__occupancy_limit_registers
dram__bytes_read.sum
dram__bytes_write.sum
dram__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct
l1tex__t_sector_hit_rate.pct
smsp__inst_executed.sum
smsp__thread_inst_executed_per_inst_executed.ratio
smsp__sass_average_branch_targets_threads_uniform.pct
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active
sm__icc_request_hit_rate.pct
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
sm__cycles_elapsed.max""".split("\n")
for k in KEYS:
    if k in col:
        print(f"{k:75s} {vals[col[k]]} {units[col[k]]}")
print("\nwarps stalled per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active.ratio):")
for h in sorted(col):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(vals[col[h]] or 0)
        if v >= 0.05:
            print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:22s} {v:.2f}")
r, w = float(vals[col["dram__bytes_read.sum"]]), float(vals[col["dram__bytes_write.sum"]])
print(f"\nDRAM traffic of the launch: {r + w:.3f} {units[col['dram__bytes_read.sum']]} (read {r:.3f} + write {w:.3f})")

# --facts NAME CAPTURE: record this launch's DRAM bytes and issue utilisation in profiles/ncu_facts.json, keyed by kernel and
# stamped with the fingerprint of the CUDA sources (bench.py quotes them only while the sources are unchanged)
if "--facts" in sys.argv:
    import json
    import os

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    import bench

    k = sys.argv.index("--facts")
    name, capture = sys.argv[k + 1], sys.argv[k + 2]
    unit = units[col["dram__bytes_read.sum"]].lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_facts.json")
    facts = json.load(open(path)) if os.path.exists(path) else {}
    facts[name] = {"dram_bytes": (r + w) * mult, "issue_active_pct": float(vals[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                   "duration_ms": float(vals[col["gpu__time_duration.sum"]]) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(
                       units[col["gpu__time_duration.sum"]].lower().replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns").replace("second", "s"), 1),
                   "capture": capture, "src": bench.source_fingerprint()}
    json.dump(facts, open(path, "w"), indent=1, sort_keys=True)
    print(f"\nrecorded {name} in {path}")
