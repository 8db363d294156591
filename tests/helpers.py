"""Shared test helpers: golden-file parsing and query construction."""
import ctypes as C
import os

import numpy as np

from riichienv_b200 import _abi as A

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_query(tiles, melds, win_tile, dora, ura, cond, player_wind, round_wind, honba):
    q = A.HandQuery()
    for i in range(14):
        q.tiles[i] = tiles[i] if i < len(tiles) else 255
    q.n_tiles = len(tiles)
    q.n_melds = len(melds)
    for mi in range(4):
        for k in range(4):
            q.meld_tiles[mi][k] = 255
    for mi, (ty, ts) in enumerate(melds):
        q.meld_type[mi] = ty
        for k in range(4):
            q.meld_tiles[mi][k] = ts[k] if k < len(ts) else 255
    q.win_tile = win_tile
    q.n_dora, q.n_ura = len(dora), len(ura)
    for i, x in enumerate(dora):
        q.dora_ind[i] = x
    for i, x in enumerate(ura):
        q.ura_ind[i] = x
    q.cond, q.player_wind, q.round_wind, q.honba = cond, player_wind, round_wind, honba
    return q


def load_agari_cases(name="agari_4p.txt"):
    """-> list of (HandQuery, (is_win, han, fu), sorted yaku ids)"""
    out = []
    for line in open(os.path.join(GOLDEN, name)):
        if line.startswith("#"):
            continue
        p = [x.strip() for x in line.split("|")]
        tiles = [int(x) for x in p[0].split(",")]
        melds = []
        for m in [m for m in p[1].split(";") if m]:
            ty, ts = m.split(":")
            melds.append((int(ty), [int(x) for x in ts.split(",")]))
        d = [int(x) for x in p[3].split(",") if x]
        u = [int(x) for x in p[4].split(",") if x]
        c = [int(x) for x in p[5].split()]
        e = tuple(int(x) for x in p[6].split())
        y = sorted(int(x) for x in p[7].split(",") if x)
        out.append((make_query(tiles, melds, int(p[2]), d, u, c[0], c[1], c[2], c[3]), e, y))
    return out


def load_counts_file(name):
    out = []
    for line in open(os.path.join(GOLDEN, name)):
        if line.startswith("#"):
            continue
        cs, v = line.split()
        out.append(([int(c) for c in cs], int(v)))
    return out


def yaku_ids(mask):
    return sorted(b for b in range(64) if (mask >> b) & 1)


def query_array(queries):
    arr = (A.HandQuery * len(queries))()
    for i, q in enumerate(queries):
        arr[i] = q
    return arr


def parse_hand(s):
    """'123m456p789s111z2z' -> list of tids (copy k of each kind in order), like riichienv.parse_hand (no red fives)."""
    cnt = [0] * 34
    digs = []
    for ch in s:
        if ch.isdigit():
            digs.append(int(ch))
        else:
            base = {"m": 0, "p": 9, "s": 18, "z": 27}[ch]
            for d in digs:
                cnt[base + d - 1] += 1
            digs = []
    tiles = []
    for t in range(34):
        for k in range(cnt[t]):
            # avoid the red-five copy (k == 0 of 5m/5p/5s) unless 4 copies are needed
            copy = k + 1 if (t in (4, 13, 22) and cnt[t] < 4) else k
            tiles.append(t * 4 + copy)
    return tiles


def random_hand_queries(n, seed):
    """Config-2 style seeded random 14-tile hands (SURVEY.md §8 d), plus structured near-complete hands."""
    rng = np.random.default_rng(seed)
    qs = []
    for i in range(n):
        if i % 2 == 0:
            tiles = rng.choice(136, 14, replace=False).tolist()
        else:
            # 4 mentsu + pair from a random multiset, then 0-2 perturbations
            cnt = [0] * 34
            groups = 0
            while groups < 4:
                if rng.random() < 0.45:
                    t = int(rng.integers(34))
                    if cnt[t] <= 1:
                        cnt[t] += 3
                        groups += 1
                else:
                    s = int(rng.integers(3)) * 9 + int(rng.integers(7))
                    if max(cnt[s:s + 3]) <= 3:
                        for k in range(3):
                            cnt[s + k] += 1
                        groups += 1
            while True:
                t = int(rng.integers(34))
                if cnt[t] <= 2:
                    cnt[t] += 2
                    break
            for _ in range(int(rng.integers(3))):
                a = [k for k in range(34) if cnt[k] > 0]
                b = [k for k in range(34) if cnt[k] < 4]
                cnt[a[int(rng.integers(len(a)))]] -= 1
                cnt[b[int(rng.integers(len(b)))]] += 1
            tiles = []
            for t in range(34):
                copies = rng.permutation(4)[: cnt[t]].tolist()
                tiles += [t * 4 + c for c in copies]
            rng.shuffle(tiles)
        win = tiles[-1]
        cond = 0
        if i & 1:
            cond |= A.C_TSUMO
        if i & 2:
            cond |= A.C_RIICHI
        if (i >> 5) & 1 and cond & A.C_RIICHI:
            cond |= A.C_IPPATSU
        dora = [int(rng.integers(136))]
        ura = [int(rng.integers(136))] if cond & A.C_RIICHI else []
        qs.append(make_query(tiles, [], win, dora, ura, cond, (i >> 2) & 3, (i >> 4) & 1, int(rng.integers(3))))
    return qs
