"""`Observation.encode_yaku_possibility` — riichienv-core/src/yaku_checker.rs (the checks) and observation/python.rs:327-449,
observation_3p/python.rs:275-395 (the (NP, 21, 2) float32 array).

A function of the observation's by-value public fields alone (every seat's melds and discards, the dora indicators, round wind,
dealer): nothing of the game record is involved, so it lives with the Observation class on the host, as in the reference.
Impossible -> 0.0; Possible and Unknown -> 1.0; both columns of a yaku carry the same value."""
POSSIBLE, IMPOSSIBLE, UNKNOWN = 1, 0, 2


def _terminal_or_honor(t):
    return t % 9 == 0 or t % 9 == 8 or t >= 27


def _visible(kind, discards, dora_indicators):          # count_visible_tiles: own discards + dora indicators
    return sum(1 for t in discards if t // 4 == kind) + sum(1 for t in dora_indicators if t // 4 == kind)


def check_tanyao(melds):
    return IMPOSSIBLE if any(_terminal_or_honor(t // 4) for m in melds for t in m.tiles) else UNKNOWN


def check_yakuhai(kind, melds, discards, dora):
    for m in melds:
        if m.tiles and m.tiles[0] // 4 == kind and len(m.tiles) >= 3:
            return POSSIBLE
    return IMPOSSIBLE if _visible(kind, discards, dora) >= 3 else UNKNOWN


def check_flush(melds):
    if not melds:
        return UNKNOWN, UNKNOWN
    kinds = [t // 4 for m in melds for t in m.tiles]
    suits = len({k // 9 for k in kinds if k < 27})
    honor = any(k >= 27 for k in kinds)
    honitsu = UNKNOWN if suits == 0 else POSSIBLE if suits == 1 else IMPOSSIBLE
    if suits == 0:
        chinitsu = UNKNOWN
    elif suits == 1 and not honor:
        chinitsu = POSSIBLE
    else:
        chinitsu = IMPOSSIBLE
    return honitsu, chinitsu


def check_toitoi(melds):
    for m in melds:
        if len(m.tiles) == 3:
            a, b, c = (t // 4 for t in m.tiles)
            if a + 1 == b and b + 1 == c and a < 27:
                return IMPOSSIBLE
    if melds and all(len(m.tiles) >= 3 and m.tiles[0] // 4 == m.tiles[1] // 4 for m in melds):
        return POSSIBLE
    return UNKNOWN


def check_chiitoitsu(melds):
    return IMPOSSIBLE if melds else UNKNOWN


def check_shousangen(melds, discards, dora):
    return IMPOSSIBLE if any(_visible(d, discards, dora) >= 4 for d in (31, 32, 33)) else UNKNOWN


def check_daisangen(melds, discards, dora):
    pons = 0
    for d in (31, 32, 33):
        if any(len(m.tiles) >= 3 and m.tiles[0] // 4 == d for m in melds):
            pons += 1
        elif _visible(d, discards, dora) >= 2:
            return IMPOSSIBLE
    return POSSIBLE if pons == 3 else UNKNOWN


def check_tsuuiisou(melds):
    return IMPOSSIBLE if any(t // 4 < 27 for m in melds for t in m.tiles) else UNKNOWN


def check_chinroutou(melds):
    for m in melds:
        for t in m.tiles:
            k = t // 4
            if k >= 27 or (k % 9 != 0 and k % 9 != 8):
                return IMPOSSIBLE
    return UNKNOWN


def check_honroutou(melds):
    for m in melds:
        for t in m.tiles:
            k = t // 4
            if k < 27 and k % 9 != 0 and k % 9 != 8:
                return IMPOSSIBLE
    return UNKNOWN


def check_kokushi(melds, discards, dora):
    if melds:
        return IMPOSSIBLE
    need = (0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33)
    return IMPOSSIBLE if any(_visible(k, discards, dora) >= 4 for k in need) else UNKNOWN


def check_chanta(melds):
    for m in melds:
        if m.tiles and not any(_terminal_or_honor(t // 4) for t in m.tiles):
            return IMPOSSIBLE
    return UNKNOWN


def check_junchan(melds):
    for m in melds:
        if not m.tiles:
            continue
        if any(t // 4 >= 27 for t in m.tiles):
            return IMPOSSIBLE
        if not any((t // 4) % 9 in (0, 8) for t in m.tiles):
            return IMPOSSIBLE
    return UNKNOWN


def check_sanshoku_doujun(melds):      # the reference is deliberately conservative: never impossible
    return UNKNOWN


def check_iipeikou(melds):
    return IMPOSSIBLE if melds else UNKNOWN


def check_ittsu(melds):                # likewise conservative
    return UNKNOWN


def encode(obs, num_players):
    """-> bytes of a (num_players, 21, 2) float32 array"""
    import numpy as np

    arr = np.ones((num_players, 21, 2), np.float32)
    dora = list(obs.dora_indicators)
    for p in range(num_players):
        melds, discards = obs.melds[p], obs.discards[p]
        seat = (p + num_players - obs.oya) % num_players
        honitsu, chinitsu = check_flush(melds)
        vals = [check_tanyao(melds),
                check_yakuhai(31, melds, discards, dora), check_yakuhai(32, melds, discards, dora), check_yakuhai(33, melds, discards, dora),
                check_yakuhai(27 + obs.round_wind, melds, discards, dora), check_yakuhai(27 + seat, melds, discards, dora),
                honitsu, chinitsu, check_toitoi(melds), check_chiitoitsu(melds), check_shousangen(melds, discards, dora),
                check_daisangen(melds, discards, dora), check_tsuuiisou(melds), check_chinroutou(melds), check_honroutou(melds),
                check_kokushi(melds, discards, dora), check_chanta(melds), check_junchan(melds), check_sanshoku_doujun(melds),
                check_iipeikou(melds), check_ittsu(melds)]
        for y, v in enumerate(vals):
            arr[p, y, :] = 0.0 if v == IMPOSSIBLE else 1.0
    return arr.tobytes()
