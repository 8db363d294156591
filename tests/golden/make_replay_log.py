#!/usr/bin/env python3
"""Writes tests/golden/126_204_0_mjai.jsonl: the event lines of the reference's own fixture tests/data/126_204_0_mjai.jsonl
(a real 12-kyoku game, 1,383 MJAI events), re-serialised one compact JSON object per line.  It is the input of the replay
ingestion tests (tests/test_replay.py): data held by the reference's tests, not source.  Run in the authoring container
(reads /root/reference); the output is what travels to the GPU box."""
import json
import os

SRC = "/root/reference/tests/data/126_204_0_mjai.jsonl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "126_204_0_mjai.jsonl")

if __name__ == "__main__":
    with open(SRC) as f, open(OUT, "w") as o:
        n = 0
        for line in f:
            if line.strip():
                o.write(json.dumps(json.loads(line), separators=(",", ":"), ensure_ascii=False) + "\n")
                n += 1
    print(f"{n} events -> {OUT}")
