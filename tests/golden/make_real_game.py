#!/usr/bin/env python3
"""Builds tests/golden/real_game_126_204_0.json from the reference's own fixture tests/data/126_204_0_mjai.jsonl — a real
12-kyoku game (1,383 MJAI events: 9 hora with `deltas` and `ura_markers`, 3 exhaustive draws, 10 riichi, 27 calls, 1 ankan +
kan dora).  Run in the authoring container (reads /root/reference); the output is what travels.

For every kyoku the walls are RECONSTRUCTED so that the environment, dealt that wall and fed the players' logged decisions,
must reproduce the log — draws, dora / ura markers, and above all the settlement numbers the reference holds as answers:
  T[0..135] is the reference's `wall.tiles` after load_wall's reverse (state/wall.rs:69-80): the deal pops from the back
  (3 x 4 tiles per seat from the dealer, then one each, state/mod.rs:1750-1765), live draws keep popping from the back,
  rinshan draws come from T[0], T[1], ..., dora indicator k = T[4 + 2k], ura k = T[5 + 2k] (state/mod.rs:2026,2051).
  Tiles the log never shows are filled in from the unused ones.
MJAI names do not identify the copy of a tile, only red fives; copies are numbered in order of appearance, which is
invisible in the event log the test compares.

Per kyoku the JSON holds: the reset() arguments, the wall as passed to reset(wall=) (i.e. reversed T), the decisions in
order (seat + the MJAI object of the log), and the expected event lines of the log (dahai / tsumo / calls / hora / ryukyoku).
"""
import json
import os
import sys

SRC = "/root/reference/tests/data/126_204_0_mjai.jsonl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "real_game_126_204_0.json")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))


def kind_of(name):
    if name in "ESWNPFC" and len(name) == 1:
        return 27 + "ESWNPFC".index(name), False
    red = name.endswith("r")
    return "mps".index(name[1]) * 9 + int(name[0]) - 1, red


class Copies:
    def __init__(self):
        self.used = set()

    def take(self, name):
        kind, red = kind_of(name)
        five = kind in (4, 13, 22)
        order = (0,) if red else ((1, 2, 3) if five else (0, 1, 2, 3))
        for c in order:
            if (kind, c) not in self.used:
                self.used.add((kind, c))
                return 4 * kind + c
        raise ValueError(f"no copy of {name} left")


def main():
    events = [json.loads(l) for l in open(SRC) if l.strip()]
    kyokus, cur = [], None
    for e in events:
        if e["type"] == "start_kyoku":
            cur = {"start": e, "events": []}
            kyokus.append(cur)
        elif e["type"] in ("start_game", "end_game"):
            continue
        elif cur is not None:
            cur["events"].append(e)
    out = []
    for ky in kyokus:
        st = ky["start"]
        oya = st["oya"]
        T = [None] * 136
        cp = Copies()
        # deal slots: d-th dealt tile sits at T[135 - d]
        d = 0
        given = {p: [cp.take(t) for t in st["tehais"][p]] for p in range(4)}
        slots = {p: [] for p in range(4)}
        for r in range(3):
            for idx in range(4):
                p = (idx + oya) % 4
                for _ in range(4):
                    slots[p].append(135 - d)
                    d += 1
        for idx in range(4):
            p = (idx + oya) % 4
            slots[p].append(135 - d)
            d += 1
        for p in range(4):
            for pos, tid in zip(slots[p], given[p]):
                T[pos] = tid
        T[4] = cp.take(st["dora_marker"])
        live, rinshan, n_dora, after_kan = 135 - 52, 0, 1, False
        decisions, expected = [], []
        for e in ky["events"]:
            ty = e["type"]
            if ty == "tsumo":
                tid = cp.take(e["pai"])
                if after_kan:
                    T[rinshan] = tid
                    rinshan += 1
                    after_kan = False
                else:
                    T[live] = tid
                    live -= 1
                expected.append(e)
            elif ty == "dora":
                T[4 + 2 * n_dora] = cp.take(e["dora_marker"])
                n_dora += 1
                expected.append(e)
            elif ty in ("dahai", "pon", "chi", "reach", "ankan", "kakan", "daiminkan", "kan"):
                decisions.append(e)
                if ty in ("ankan", "kakan", "daiminkan", "kan"):
                    after_kan = True
                expected.append(e)
            elif ty == "hora":
                for k, name in enumerate(e.get("ura_markers", [])):
                    if T[5 + 2 * k] is None:
                        T[5 + 2 * k] = cp.take(name)
                decisions.append(e)
                expected.append(e)
            elif ty in ("reach_accepted", "ryukyoku"):
                expected.append(e)
        rest = [t for t in range(136) if (t // 4, t % 4) not in cp.used]
        for i in range(136):
            if T[i] is None:
                T[i] = rest.pop()
        assert sorted(T) == list(range(136))
        out.append({"oya": oya, "bakaze": "ESWN".index(st["bakaze"]), "honba": st["honba"], "kyotaku": st["kyotaku"],
                    "scores": st["scores"], "wall": T[::-1], "dora_marker": st["dora_marker"],
                    "decisions": decisions, "expected": expected})
    json.dump({"source": "tests/data/126_204_0_mjai.jsonl of smly/RiichiEnv (reformatted: walls reconstructed per kyoku)",
               "kyokus": out}, open(OUT, "w"), separators=(",", ":"))
    print(f"{len(out)} kyoku, {sum(len(k['expected']) for k in out)} expected events -> {OUT} ({os.path.getsize(OUT)} bytes)")


if __name__ == "__main__":
    main()
