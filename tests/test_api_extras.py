"""Small pieces of the reference's public surface next to the hand evaluator: the yaku catalogue (yaku.rs:11-127),
WinResult.yaku_list (types.rs:356-361), check_riichi_candidates (hand_evaluator.rs:263-284), consts."""
import random

import pytest


def _shim(backend):
    from tests.test_real_game_replay import _env_module

    return _env_module(backend)


def _agari(counts):
    """plain backtracking agari test on a 34-histogram (standard form, seven pairs, thirteen orphans)"""
    n = sum(counts)
    if n == 14:
        if all(c in (0, 2) for c in counts):
            return True
        yao = [0, 8, 9, 17, 18, 26, 27, 28, 29, 30, 31, 32, 33]
        if all(counts[k] >= 1 for k in yao) and sum(counts[k] for k in yao) == 14:
            return True

    def sets(c):
        i = next((k for k in range(34) if c[k]), None)
        if i is None:
            return True
        if c[i] >= 3:
            c[i] -= 3
            ok = sets(c)
            c[i] += 3
            if ok:
                return True
        if i < 27 and i % 9 <= 6 and c[i + 1] and c[i + 2]:
            for k in (i, i + 1, i + 2):
                c[k] -= 1
            ok = sets(c)
            for k in (i, i + 1, i + 2):
                c[k] += 1
            if ok:
                return True
        return False

    for p in range(34):
        if counts[p] >= 2:
            counts[p] -= 2
            ok = sets(counts)
            counts[p] += 2
            if ok:
                return True
    return False


def _candidates_brute(tiles):
    out = []
    for i, t in enumerate(tiles):
        c = [0] * 34
        for j, u in enumerate(tiles):
            if j != i:
                c[u // 4] += 1
        for k in range(34):
            if c[k] < 4:
                c[k] += 1
                ok = _agari(c)
                c[k] -= 1
                if ok:
                    out.append(t)
                    break
    return out


def _near_hands(rng, n_tiles):
    """a complete hand of n_tiles (sets + pair) with one or two tiles swapped for random ones: mostly tenpai / one away"""
    while True:
        c = [0] * 34
        tiles = []
        def take(k):
            if c[k] >= 4:
                return False
            tiles.append(4 * k + c[k])
            c[k] += 1
            return True
        ok = True
        for _ in range(n_tiles // 3):
            if rng.random() < 0.6:
                s = rng.randrange(3) * 9 + rng.randrange(7)
                ok &= all(take(s + d) for d in range(3))
            else:
                k = rng.randrange(34)
                ok &= all(take(k) for _ in range(3))
        k = rng.randrange(34)
        ok &= take(k) and take(k)
        if not ok:
            continue
        for _ in range(rng.randrange(0, 3)):
            i = rng.randrange(len(tiles))
            pool = [t for t in range(136) if t not in tiles]
            tiles[i] = rng.choice(pool)
        rng.shuffle(tiles)
        return tiles


@pytest.mark.parametrize("backend", ["oracle", "hostsim"])
def test_check_riichi_candidates_equals_brute_force(backend):
    m = _shim(backend)
    rng = random.Random(5)
    hits = 0
    for n_tiles in (14, 14, 14, 11, 8, 5, 2):
        for _ in range(40):
            tiles = _near_hands(rng, n_tiles)
            got = m.check_riichi_candidates(tiles)
            assert got == _candidates_brute(tiles), tiles
            hits += bool(got)
    assert hits > 100
    # seven pairs and thirteen orphans count (agari.rs:23-30), input order and duplicates of a kind are kept
    chiitoi = [0, 1, 8, 9, 16, 17, 40, 41, 72, 73, 108, 109, 120, 132]
    assert m.check_riichi_candidates(chiitoi) == [120, 132]
    kokushi = [0, 32, 36, 68, 72, 104, 108, 112, 116, 120, 124, 128, 132, 60]
    assert m.check_riichi_candidates(kokushi) == [60]
    assert m.check_riichi_candidates([]) == []
    with pytest.raises(ValueError):
        m.check_riichi_candidates(list(range(13)))


def test_yaku_catalogue_and_yaku_list():
    m = _shim("oracle")
    ys = m.get_all_yaku()
    assert len(ys) == 49 and [y.id for y in ys] == sorted(y.id for y in ys) and all(y.id == y.mjsoul_id for y in ys)
    assert m.get_yaku_by_id(46) is None and m.get_yaku_by_id(99) is None
    r = m.get_yaku_by_id(2)
    assert (r.name, r.name_en, r.tenhou_id) == ("立直", "Riichi", 1)
    assert repr(r) == "Yaku(id=2, name='立直', name_en='Riichi', tenhou_id=1, mjsoul_id=2)"
    assert len({y.tenhou_id for y in ys}) == 48            # dora and nukidora share tenhou id 52
    # the reference's README example (README.md:218-223): WinResult(... yaku=[8, 11, 10, 22], han=5, fu=60)
    res = m.HandEvaluator.hand_from_text("111m33p12s111666z").calc(m.convert.mpsz_to_tid("3s"), dora_indicators=[], ura_indicators=[])
    assert (res.is_win, res.ron_agari, res.yaku, res.han, res.fu) == (True, 12000, [8, 11, 10, 22], 5, 60)
    assert [y.id for y in res.yaku_list()] == res.yaku
    assert [y.name_en for y in res.yaku_list()] == ["Yakuhai (hatsu)", "Yakuhai (round wind)", "Yakuhai (seat wind)", "San Ankou"]
    # every id the evaluator can report is in the catalogue
    from riichienv_b200.hand import _ORDER_CHIITOI, _ORDER_STANDARD
    assert all(m.get_yaku_by_id(i) is not None for i in set(_ORDER_STANDARD) | set(_ORDER_CHIITOI))
    assert (m.consts.N_TILE_TYPES_4P, m.consts.N_TILE_TYPES_3P, m.consts.N_TILES_4P, m.consts.N_TILES_3P) == (34, 27, 136, 108)
    assert m.WinResultContext is not None and m.WinResultContextIterator is not None


@pytest.mark.gpu
def test_gpu_check_riichi_candidates():
    import riichienv_b200 as rb

    rng = random.Random(6)
    for n_tiles in (14, 14, 11, 8, 5, 2):
        for _ in range(10):
            tiles = _near_hands(rng, n_tiles)
            assert rb.check_riichi_candidates(tiles) == _candidates_brute(tiles), tiles
