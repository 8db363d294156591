#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the reference checkout.

Run in the authoring container only (it reads /root/reference, which does not
exist on the GPU box).  The outputs are committed; tests never read the
reference at run time.

Sources (relative to /root/reference/riichienv-core):
  benches/data/agari_4p.json   816 winning 4P hands  (tests/agari_correctness.rs:137-199)
  benches/data/agari_3p.json   402 winning 3P hands  (tests/agari_correctness.rs:201-262)
  benches/data/hands_negative.json  200 non-agari 34-histograms with is_tenpai
  src/data/nyanten_*.bin + src/shanten.rs:6-239  -> shanten of seeded random hands,
      computed here by a Python restatement of the reference's table walk, so the
      numbers are the reference tables' own answers.
"""
import json
import os
import random
import re
import sys

REF = "/root/reference/riichienv-core"
OUT = os.path.dirname(os.path.abspath(__file__))

MELD = {"chi": 0, "pon": 1, "daiminkan": 2, "ankan": 3, "kakan": 4}
COND_BITS = [
    ("tsumo", 0x001), ("riichi", 0x002), ("double_riichi", 0x004), ("ippatsu", 0x008), ("haitei", 0x010),
    ("houtei", 0x020), ("rinshan", 0x040), ("chankan", 0x080), ("tsumo_first_turn", 0x100),
]


def conv_agari(src, dst):
    cases = json.load(open(os.path.join(REF, "benches/data", src)))["cases"]
    with open(os.path.join(OUT, dst), "w") as f:
        f.write("# tiles | melds(type:tiles;...) | win | dora | ura | cond pw rw honba kita sanma np | is_win han fu | yaku\n")
        for c in cases:
            cd = c["conditions"]
            bits = sum(b for k, b in COND_BITS if cd[k])
            melds = ";".join(f"{MELD[m['meld_type']]}:{','.join(map(str, m['tiles']))}" for m in c["melds"])
            e = c["expected"]
            f.write(" | ".join([
                ",".join(map(str, c["tiles_136"])), melds, str(c["win_tile_136"]),
                ",".join(map(str, c["dora_indicators"])), ",".join(map(str, c["ura_indicators"])),
                f"{bits} {cd['player_wind']} {cd['round_wind']} {cd['honba']} {cd.get('kita_count', 0)} "
                f"{int(cd.get('is_sanma', False))} {cd.get('num_players', 4)}",
                f"{int(e['is_win'])} {e['han']} {e['fu']}", ",".join(map(str, e["yaku"])),
            ]) + "\n")
    print(dst, len(cases))


def conv_negative():
    cases = json.load(open(os.path.join(REF, "benches/data/hands_negative.json")))["cases"]
    with open(os.path.join(OUT, "hands_negative.txt"), "w") as f:
        f.write("# counts_34 (34 digits) is_tenpai\n")
        for c in cases:
            f.write("".join(map(str, c["counts_34"])) + f" {int(c['is_tenpai'])}\n")
    print("hands_negative.txt", len(cases))


def load_shanten_tables():
    src = open(os.path.join(REF, "src/shanten.rs")).read()

    def table(name):
        m = re.search(name + r": \[\[\[u32; 5\]; (\d+)\]; (\d+)\] = \[(.*?)\n\];", src, re.S)
        nums = list(map(int, re.findall(r"\d+", re.sub(r"//[^\n]*", "", m.group(3)))))
        n1, n0 = int(m.group(1)), int(m.group(2))
        assert len(nums) == n0 * n1 * 5, (name, len(nums))
        return [[nums[(i * n1 + j) * 5:(i * n1 + j) * 5 + 5] for j in range(n1)] for i in range(n0)]

    d = os.path.join(REF, "src/data")
    rd = lambda n: open(os.path.join(d, n), "rb").read()
    return dict(
        shupai=table("SHUPAI_TABLE"), zipai=table("ZIPAI_TABLE"),
        sk=rd("nyanten_shupai_keys.bin"), zk=rd("nyanten_zipai_keys.bin"),
        k1=rd("nyanten_keys1.bin"), k2=rd("nyanten_keys2.bin"), k3=rd("nyanten_keys3.bin"),
    )


def ref_shanten(T, cnt, m):
    """shanten.rs:163-239 restated."""
    def h(tab, tiles):
        n = 0
        hv = 0
        for i, c in enumerate(tiles):
            n += c
            hv += tab[i][n][c]
        return hv
    k0m = T["sk"][h(T["shupai"], cnt[0:9])]
    k0p = T["sk"][h(T["shupai"], cnt[9:18])]
    k1 = T["k1"][k0m * 126 + k0p]
    k0s = T["sk"][h(T["shupai"], cnt[18:27])]
    k2 = T["k2"][k1 * 126 + k0s]
    k0z = T["zk"][h(T["zipai"], cnt[27:34])]
    s = T["k3"][(k2 * 55 + k0z) * 5 + m] - 1
    if s <= 0 or m < 4:
        return s
    kinds = sum(1 for c in cnt if c > 0)
    pairs = sum(1 for c in cnt if c >= 2)
    s = min(s, 7 - pairs + max(0, 7 - kinds) - 1)
    if s > 0:
        term = [0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33]
        k = sum(1 for i in term if cnt[i] > 0)
        p = any(cnt[i] >= 2 for i in term)
        s = min(s, 14 - k - int(p) - 1)
    return s


def gen_shanten():
    T = load_shanten_tables()
    rng = random.Random(20251017)
    lines = []
    # uniformly random hands of 1..14 tiles, plus "structured" hands biased to low shanten
    for it in range(6000):
        n = rng.choice([13, 14, 13, 14, 13, 14, 10, 11, 7, 8, 4, 5, 1, 2])
        if it % 3 == 0:
            tiles = rng.sample(range(136), n)
        else:
            # build from mentsu/pairs then perturb: gives shanten -1..2 coverage
            cnt = [0] * 34
            left = n
            while left >= 3:
                if rng.random() < 0.5:
                    t = rng.randrange(34)
                    if cnt[t] <= 1:
                        cnt[t] += 3
                        left -= 3
                else:
                    s = rng.randrange(3) * 9 + rng.randrange(7)
                    if max(cnt[s:s + 3]) <= 3:
                        for k in range(3):
                            cnt[s + k] += 1
                        left -= 3
            while left > 0:
                t = rng.randrange(34)
                if cnt[t] < 4:
                    cnt[t] += 1
                    left -= 1
            for _ in range(rng.randrange(3)):
                a = rng.choice([i for i in range(34) if cnt[i] > 0])
                b = rng.choice([i for i in range(34) if cnt[i] < 4])
                cnt[a] -= 1
                cnt[b] += 1
            tiles = [t * 4 + k for t in range(34) for k in range(cnt[t])]
        cnt = [0] * 34
        for t in tiles:
            cnt[t // 4] += 1
        s = ref_shanten(T, cnt, len(tiles) // 3)
        lines.append("".join(map(str, cnt)) + f" {s}\n")
    with open(os.path.join(OUT, "shanten_golden.txt"), "w") as f:
        f.write("# counts_34 (34 digits) shanten   [reference tables' answer, len_div3 = n_tiles // 3]\n")
        f.writelines(lines)
    print("shanten_golden.txt", len(lines))
    # sanity: README / tests/test_shanten.py known answers
    def parse(s):
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
        return cnt
    for hs, exp in [("19m19p19s1234567z", 0), ("1199m1199p1199s1z", 0), ("111m999m111p11z", -1),
                    ("19m147p258s12345z", 5), ("1111m111122233z", 1), ("11119999m22345s", 1)]:
        c = parse(hs)
        assert ref_shanten(T, c, sum(c) // 3) == exp, hs


def structured_hand(rng, n):
    """n tiles built from mentsu/pairs then perturbed (shanten -1..2 coverage)."""
    cnt = [0] * 34
    left = n
    while left >= 3:
        if rng.random() < 0.5:
            t = rng.randrange(34)
            if cnt[t] <= 1:
                cnt[t] += 3
                left -= 3
        else:
            s = rng.randrange(3) * 9 + rng.randrange(7)
            if max(cnt[s:s + 3]) <= 3:
                for k in range(3):
                    cnt[s + k] += 1
                left -= 3
    while left > 0:
        t = rng.randrange(34)
        if cnt[t] < 4:
            cnt[t] += 1
            left -= 1
    for _ in range(rng.randrange(3)):
        a = rng.choice([i for i in range(34) if cnt[i] > 0])
        b = rng.choice([i for i in range(34) if cnt[i] < 4])
        cnt[a] -= 1
        cnt[b] += 1
    return cnt


def gen_ukeire():
    """shanten.rs:250-393 (calculate_shanten / calculate_effective_tiles_with_discard / calculate_best_ukeire) restated on
    top of the reference's own tables: hands of 13/14 (closed), 10/11, 7/8, 4/5 tiles and a random visible-tile list."""
    T = load_shanten_tables()
    rng = random.Random(20261017)

    def sh(tiles):
        cnt = [0] * 34
        for t in tiles:
            cnt[t // 4] += 1
        return ref_shanten(T, cnt, len(tiles) // 3)

    def effective(hand):
        cur = sh(hand)
        hc = [0] * 34
        for t in hand:
            hc[t // 4] += 1
        return sum(1 for k in range(34) if hc[k] < 4 and sh(hand + [k * 4]) < cur)

    def effective_with_discard(hand):
        if len(hand) % 3 == 1:
            return effective(hand)
        s0 = sh(hand)
        best = 0
        for i in range(len(hand)):
            sub = hand[:i] + hand[i + 1:]
            if sh(sub) <= s0:
                best = max(best, effective(sub))
        return best

    def best_ukeire(hand, visible):
        vis = [0] * 34
        for t in visible:
            vis[t // 4] += 1
        cur = sh(hand)
        base = [0] * 34
        for t in hand:
            base[t // 4] += 1
        best = 0
        for i in range(len(hand)):
            sub = hand[:i] + hand[i + 1:]
            nc = list(base)
            nc[hand[i] // 4] -= 1
            ns = sh(sub)
            if ns > cur:
                continue
            uke = 0
            for k in range(34):
                if nc[k] >= 4:
                    continue
                if sh(sub + [k * 4]) < ns:
                    uke += max(0, max(0, 4 - vis[k]) - nc[k])
            best = max(best, uke)
        return best

    lines = []
    for it in range(400):
        n = rng.choice([13, 14, 13, 14, 13, 14, 10, 11, 7, 8, 4, 5])
        if it % 4 == 0:
            hand = sorted(rng.sample(range(136), n))
        else:
            cnt = structured_hand(rng, n)
            hand = [t * 4 + k for t in range(34) for k in range(cnt[t])]
        rest = [t for t in range(136) if t not in hand]
        visible = rng.sample(rest, rng.randrange(0, 60))
        lines.append(",".join(map(str, hand)) + " | " + ",".join(map(str, visible)) +
                     f" | {sh(hand)} {effective_with_discard(hand)} {best_ukeire(hand, visible)}\n")
    with open(os.path.join(OUT, "ukeire_golden.txt"), "w") as f:
        f.write("# hand tids | visible tids | shanten effective_tiles_with_discard best_ukeire   [reference tables' answers]\n")
        f.writelines(lines)
    print("ukeire_golden.txt", len(lines))


def ref_shanten_3p(T, cnt, m):
    """shanten.rs:407-468 restated (relocation of 1m / 9m into empty honor slots, chiitoi without 2m-8m)."""
    t = list(cnt)
    mc = [t[0], t[8]]
    t[0] = t[8] = 0
    slot = 27
    for i in range(2):
        if mc[i] == 0:
            continue
        while slot < 34 and t[slot] != 0:
            slot += 1
        if slot < 34:
            t[slot] = mc[i]
            slot += 1
        else:
            t[(0, 8)[i]] = mc[i]
    # normal form through the 4P table walk (without its chiitoi / kokushi closing)
    def h(tab, tiles):
        n = 0
        hv = 0
        for i, c in enumerate(tiles):
            n += c
            hv += tab[i][n][c]
        return hv
    k0m = T["sk"][h(T["shupai"], t[0:9])]
    k0p = T["sk"][h(T["shupai"], t[9:18])]
    k1 = T["k1"][k0m * 126 + k0p]
    k0s = T["sk"][h(T["shupai"], t[18:27])]
    k2 = T["k2"][k1 * 126 + k0s]
    k0z = T["zk"][h(T["zipai"], t[27:34])]
    s = T["k3"][(k2 * 55 + k0z) * 5 + m] - 1
    if s <= 0 or m < 4:
        return s
    kinds = sum(1 for i, c in enumerate(cnt) if c > 0 and not 1 <= i <= 7)
    pairs = sum(1 for i, c in enumerate(cnt) if c >= 2 and not 1 <= i <= 7)
    s = min(s, 7 - pairs + max(0, 7 - kinds) - 1)
    if s > 0:
        term = [0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33]
        k = sum(1 for i in term if cnt[i] > 0)
        p = any(cnt[i] >= 2 for i in term)
        s = min(s, 14 - k - int(p) - 1)
    return s


def gen_shanten_3p():
    T = load_shanten_tables()
    rng = random.Random(20261018)
    valid = [0] + list(range(8, 34))
    lines = []
    for it in range(3000):
        n = rng.choice([13, 14, 13, 14, 13, 14, 10, 11, 7, 8, 4, 5, 12, 9])
        cnt = [0] * 34
        if it % 10 == 0:
            # overflow of the relocation: every honor present, plus 1m and 9m
            for k in range(27, 34):
                cnt[k] = 1
            cnt[0] = rng.randrange(1, 4)
            cnt[8] = rng.randrange(1, 4)
            left = max(0, 14 - sum(cnt))
        elif it % 10 == 1:
            left = n                      # arbitrary tids, 2m-8m included (the functions accept them)
            valid_here = list(range(34))
        else:
            left = n
        pool = valid_here if it % 10 == 1 else valid
        if it % 3 == 0 or it % 10 <= 1:
            while left > 0:
                t = rng.choice(pool)
                if cnt[t] < 4:
                    cnt[t] += 1
                    left -= 1
        else:
            while left >= 3:
                if rng.random() < 0.5:
                    t = rng.choice(pool)
                    if cnt[t] <= 1:
                        cnt[t] += 3
                        left -= 3
                else:
                    s0 = rng.choice([9, 18]) + rng.randrange(7)
                    if max(cnt[s0:s0 + 3]) <= 3:
                        for k in range(3):
                            cnt[s0 + k] += 1
                        left -= 3
            while left > 0:
                t = rng.choice(pool)
                if cnt[t] < 4:
                    cnt[t] += 1
                    left -= 1
            for _ in range(rng.randrange(3)):
                a = rng.choice([i for i in range(34) if cnt[i] > 0])
                b = rng.choice([i for i in pool if cnt[i] < 4])
                cnt[a] -= 1
                cnt[b] += 1
        s = ref_shanten_3p(T, cnt, sum(cnt) // 3)
        lines.append("".join(map(str, cnt)) + f" {s}\n")
    with open(os.path.join(OUT, "shanten3p_golden.txt"), "w") as f:
        f.write("# counts_34 (34 digits) shanten_3p   [reference tables' answer through calc_shanten_from_counts_3p, len_div3 = n // 3]\n")
        f.writelines(lines)
    print("shanten3p_golden.txt", len(lines))
    # tests/test_shanten.py known answers for calculate_shanten_3p


def gen_ukeire_3p():
    """shanten.rs:470-615 (calculate_shanten_3p / calculate_effective_tiles_3p_with_discard / calculate_best_ukeire_3p)
    restated on top of the reference's own tables."""
    T = load_shanten_tables()
    rng = random.Random(20261019)
    valid = [0] + list(range(8, 34))

    def sh(tiles):
        cnt = [0] * 34
        for t in tiles:
            cnt[t // 4] += 1
        return ref_shanten_3p(T, cnt, len(tiles) // 3)

    def effective(hand):
        cur = sh(hand)
        hc = [0] * 34
        for t in hand:
            hc[t // 4] += 1
        return sum(1 for k in valid if hc[k] < 4 and sh(hand + [k * 4]) < cur)

    def effective_with_discard(hand):
        if len(hand) % 3 == 1:
            return effective(hand)
        s0 = sh(hand)
        best = 0
        for i in range(len(hand)):
            sub = hand[:i] + hand[i + 1:]
            if sh(sub) <= s0:
                best = max(best, effective(sub))
        return best

    def best_ukeire(hand, visible):
        vis = [0] * 34
        for t in visible:
            vis[t // 4] += 1
        cur = sh(hand)
        base = [0] * 34
        for t in hand:
            base[t // 4] += 1
        best = 0
        for i in range(len(hand)):
            sub = hand[:i] + hand[i + 1:]
            nc = list(base)
            nc[hand[i] // 4] -= 1
            ns = sh(sub)
            if ns > cur:
                continue
            uke = 0
            for k in valid:
                if nc[k] >= 4:
                    continue
                if sh(sub + [k * 4]) < ns:
                    uke += max(0, max(0, 4 - vis[k]) - nc[k])
            best = max(best, uke)
        return best

    tiles108 = [t for t in range(136) if (t // 4) in valid]
    lines = []
    for it in range(300):
        n = rng.choice([13, 14, 13, 14, 13, 14, 10, 11, 7, 8, 4, 5])
        if it % 4 == 0:
            hand = sorted(rng.sample(tiles108, n))
        else:
            cnt = [0] * 34
            left = n
            while left >= 3:
                if rng.random() < 0.5:
                    t = rng.choice(valid)
                    if cnt[t] <= 1:
                        cnt[t] += 3
                        left -= 3
                else:
                    s0 = rng.choice([9, 18]) + rng.randrange(7)
                    if max(cnt[s0:s0 + 3]) <= 3:
                        for k in range(3):
                            cnt[s0 + k] += 1
                        left -= 3
            while left > 0:
                t = rng.choice(valid)
                if cnt[t] < 4:
                    cnt[t] += 1
                    left -= 1
            hand = [t * 4 + k for t in range(34) for k in range(cnt[t])]
        rest = [t for t in tiles108 if t not in hand]
        visible = rng.sample(rest, rng.randrange(0, 50))
        lines.append(",".join(map(str, hand)) + " | " + ",".join(map(str, visible)) +
                     f" | {sh(hand)} {effective_with_discard(hand)} {best_ukeire(hand, visible)}\n")
    with open(os.path.join(OUT, "ukeire3p_golden.txt"), "w") as f:
        f.write("# hand tids | visible tids | shanten_3p effective_tiles_3p_with_discard best_ukeire_3p   [reference tables' answers]\n")
        f.writelines(lines)
    print("ukeire3p_golden.txt", len(lines))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference checkout not found: fixtures are committed, nothing to do")
    conv_agari("agari_4p.json", "agari_4p.txt")
    conv_agari("agari_3p.json", "agari_3p.txt")
    conv_negative()
    gen_shanten()
    gen_ukeire()
    gen_shanten_3p()
    gen_ukeire_3p()
