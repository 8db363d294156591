"""Pins the ORACLE (CPU restatement) on the reference's own golden vectors / known answers.

Sources: riichienv-core/benches/data/*.json (via tests/golden/*.txt),
riichienv-core/tests/agari_correctness.rs:286-348 (score table),
tests/test_agari_calculator.py:41-97, README.md:221-223, tests/test_shanten.py.
"""
import ctypes as C

import pytest

import oracle
from riichienv_b200 import _abi as A
from tests import helpers as H


@pytest.fixture(scope="module")
def orc():
    return oracle.load()


def test_agari_4p_golden(orc):
    cases = H.load_agari_cases()
    assert len(cases) == 816
    for i, (q, exp, yaku) in enumerate(cases):
        r = A.HandResult()
        orc.orc_hand_eval(C.byref(q), C.byref(r), 1)
        assert (r.is_win, r.han, r.fu) == exp, f"case {i}"
        assert H.yaku_ids(r.yaku_mask) == yaku, f"case {i}"
        assert r.has_win_shape == 1


def test_agari_3p_golden(orc):
    """402 sanma cases (benches/data/agari_3p.json; tests/agari_correctness.rs:201-262) through HandEvaluator3P semantics."""
    import os

    cases = H.load_agari_cases("agari_3p.txt")
    lines = [l for l in open(os.path.join(H.GOLDEN, "agari_3p.txt")) if not l.startswith("#")]
    assert len(cases) == 402
    for i, ((q, exp, yaku), line) in enumerate(zip(cases, lines)):
        c = [int(x) for x in line.split("|")[5].split()]
        q.sanma, q.kita_count = 1, c[4]
        r = A.HandResult()
        orc.orc_hand_eval(C.byref(q), C.byref(r), 1)
        assert (r.is_win, r.han, r.fu) == exp and H.yaku_ids(r.yaku_mask) == yaku, f"case {i}"


def test_negative_hands(orc):
    cases = H.load_counts_file("hands_negative.txt")
    assert len(cases) == 200
    for cnt, tenpai in cases:
        arr = (C.c_uint8 * 34)(*cnt)
        assert orc.orc_is_agari(arr) == 0
        if sum(cnt) == 13:
            assert orc.orc_is_tenpai_counts(arr) == tenpai


def test_shanten_golden(orc):
    cases = H.load_counts_file("shanten_golden.txt")
    assert len(cases) == 6000
    for cnt, sh in cases:
        assert orc.orc_shanten_counts((C.c_uint8 * 34)(*cnt), sum(cnt) // 3) == sh


SHANTEN_KNOWN = [  # tests/test_shanten.py
    ("1111m111122233z", 1), ("111m111z222z333z44z", -1), ("123456789p11222z", -1), ("111m123456789s11z", -1),
    ("19m19p19s1234567z", 0), ("111m999m123p789s1z", 0), ("1199m1199p1199s1z", 0), ("11m99m123p456s111z", 0),
    ("111m999m123p13s7z", 1), ("11119999m22345s", 1), ("1111m9m1234567z", 3), ("111m999m111p11z", -1),
    ("111m123456789p1z", 0), ("999m111222333z1p", 0), ("11m99m11p99p11s99s1z", 0), ("111999m111999p1z", 0),
    ("19m147p258s12345z", 5),
]


@pytest.mark.parametrize("hand,expected", SHANTEN_KNOWN)
def test_shanten_known_answers(orc, hand, expected):
    cnt = [0] * 34
    for t in H.parse_hand(hand):
        cnt[t // 4] += 1
    assert orc.orc_shanten_counts((C.c_uint8 * 34)(*cnt), sum(cnt) // 3) == expected


SCORE_TABLE = [  # tests/agari_correctness.rs:290-331
    (1, 30, 0, 0, 0, 4, 1000, 0, 0), (1, 30, 0, 1, 0, 4, 0, 500, 300), (3, 30, 1, 0, 0, 4, 5800, 0, 0),
    (5, 0, 0, 0, 0, 4, 8000, 0, 0), (5, 0, 1, 1, 0, 4, 0, 0, 4000), (5, 0, 0, 1, 0, 4, 0, 4000, 2000),
    (6, 0, 0, 0, 0, 4, 12000, 0, 0), (8, 0, 0, 0, 0, 4, 16000, 0, 0), (11, 0, 0, 0, 0, 4, 24000, 0, 0),
    (13, 0, 0, 0, 0, 4, 32000, 0, 0), (13, 0, 1, 0, 0, 4, 48000, 0, 0), (13, 0, 0, 1, 0, 4, 0, 16000, 8000),
    (13, 0, 1, 1, 0, 4, 0, 0, 16000), (26, 0, 0, 0, 0, 4, 64000, 0, 0), (26, 0, 1, 0, 0, 4, 96000, 0, 0),
    (26, 0, 0, 1, 0, 4, 0, 32000, 16000), (26, 0, 1, 1, 0, 4, 0, 0, 32000), (26, 0, 0, 0, 2, 4, 64600, 0, 0),
    (26, 0, 1, 1, 2, 4, 0, 200, 32200), (39, 0, 0, 0, 0, 4, 96000, 0, 0), (39, 0, 1, 1, 0, 4, 0, 0, 48000),
    (52, 0, 0, 0, 0, 4, 128000, 0, 0), (65, 0, 0, 0, 0, 4, 160000, 0, 0), (13, 0, 0, 0, 0, 3, 32000, 0, 0),
    (13, 0, 0, 1, 0, 3, 0, 16000, 8000), (26, 0, 0, 0, 0, 3, 64000, 0, 0), (26, 0, 1, 1, 0, 3, 0, 0, 32000),
]


@pytest.mark.parametrize("case", SCORE_TABLE)
def test_score_table(orc, case):
    han, fu, oya, tsumo, honba, np_, ron, p_oya, p_ko = case
    out = (C.c_uint32 * 4)()
    orc.orc_calculate_score(han, fu, oya, tsumo, honba, np_, out)
    assert (out[0], out[1], out[2]) == (ron, p_oya, p_ko)


def _calc(orc, hand, win, cond=0, pw=0, rw=0, melds=(), dora=()):
    q = H.make_query(H.parse_hand(hand), list(melds), win, list(dora), [], cond, pw, rw, 0)
    r = A.HandResult()
    orc.orc_hand_eval(C.byref(q), C.byref(r), 1)
    return r


def test_readme_example(orc):
    # README.md:221-223: 111m33p12s111666z + 3s ron -> 12000 for dealer?  yaku [8, 11, 10, 22], 5 han 60 fu
    # (hatsu triplet 666z is kind 32 -> id 8; 111z is East: round + seat wind)
    tiles = H.parse_hand("111m33p12s111666z")
    q = H.make_query(tiles, [], 18 * 4 + 2 * 4, [], [], 0, 0, 0, 0)  # 3s
    r = A.HandResult()
    orc.orc_hand_eval(C.byref(q), C.byref(r), 1)
    assert r.is_win and r.han == 5 and r.fu == 60
    assert H.yaku_ids(r.yaku_mask) == [8, 10, 11, 22]
    assert r.ron_agari == 12000


def test_agari_calculator_known(orc):
    # tests/test_agari_calculator.py:41-97: 123m456p789s111z + 2z pair wait, winds vary the yakuhai count
    tiles = H.parse_hand("123m456p789s111z2z")
    win = 28 * 4 + 1
    # seat East, round East: double-wind triplet -> 2 han
    q = H.make_query(tiles, [], win, [], [], 0, 0, 0, 0)
    r = A.HandResult()
    orc.orc_hand_eval(C.byref(q), C.byref(r), 1)
    assert r.is_win and r.han == 2 and r.fu == 40
    # seat South, round South: 111z is no yakuhai -> no yaku on ron
    q = H.make_query(tiles, [], win, [], [], 0, 1, 1, 0)
    orc.orc_hand_eval(C.byref(q), C.byref(r), 1)
    assert r.has_win_shape and not r.is_win


def test_kokushi_and_chiitoi(orc):
    r = _calc(orc, "19m19p19s1234567z", 0 * 4 + 1)  # 13-sided
    assert r.is_win and r.yakuman and r.han == 26 and H.yaku_ids(r.yaku_mask) == [49]
    r = _calc(orc, "1122m3344p5566s7z", 33 * 4 + 1)
    assert r.is_win and r.fu == 25 and 25 in H.yaku_ids(r.yaku_mask)
    # ryanpeikou shape is never scored as chiitoitsu (yaku.rs:236-296)
    r = _calc(orc, "112233m445566p7s", 24 * 4 + 1, cond=A.C_RIICHI)
    assert r.is_win and 28 in H.yaku_ids(r.yaku_mask) and 25 not in H.yaku_ids(r.yaku_mask)


def test_sequence_feature_known_answers():
    """Known answers of the reference's own unit tests (observation/sequence_features.rs:845-930)."""
    o = oracle.load()
    # test_tile_id_to_kan37
    for tid, k in ((16, 0), (52, 10), (88, 20), (0, 1), (3, 1), (17, 5), (32, 9), (36, 11), (72, 21), (108, 30), (132, 36)):
        assert o.orc_seq_kan37(tid) == k
    # test_relative_from
    for (a, t), r in (((0, 3), 2), ((0, 1), 0), ((0, 2), 1), ((2, 3), 0)):
        assert o.orc_seq_relative_from(a, t) == r
    # test_encode_chi_basic
    assert o.orc_seq_encode_chi(4, 8, 0) == 0
    assert o.orc_seq_encode_chi(0, 8, 4) == 1
    # test_encode_pon_honor / test_encode_pon_five_red
    assert o.orc_seq_encode_pon(109, 110, 108) == 33
    assert o.orc_seq_encode_pon(16, 17, 18) == 5
    assert o.orc_seq_encode_pon(17, 18, 16) == 6
    # bounds (test_progression_type_bounds / test_candidate_type_bounds): chi 0..89, pon 0..39 over every legal call
    chis, pons = set(), set()
    for suit in range(3):
        for start in range(7):
            kinds = [suit * 9 + start + d for d in range(3)]
            for call in range(3):
                for copies in ((0, 0, 0), (1, 1, 1)):
                    t = [4 * k + c for k, c in zip(kinds, copies)]
                    rest = [x for i, x in enumerate(t) if i != call]
                    chis.add(o.orc_seq_encode_chi(rest[0], rest[1], t[call]))
    for kind in range(34):
        for c in ((0, 1, 2), (1, 2, 0), (1, 2, 3)):
            pons.add(o.orc_seq_encode_pon(4 * kind + c[0], 4 * kind + c[1], 4 * kind + c[2]))
    assert min(chis) == 0 and max(chis) == 89 and len(chis) == 90
    assert min(pons) == 0 and max(pons) == 39 and len(pons) == 40


def test_chacha_known_answers():
    """Partial pin of the seeded-wall boundary (state/wall.rs:38-48 -> crates rand 0.10 / chacha20 0.10, not vendored).

    The ChaCha block function, state layout (constants | key | 64-bit counter | stream 0) and the little-endian word
    order the wall shuffle consumes are checked against published vectors:
      * ChaCha20, zero key/nonce, block 0 (the IETF / djb keystream that rand_chacha's `test_chacha_true_values_a` holds),
      * ChaCha12, zero key/nonce, block 0 (draft-strombergson-chacha-test-vectors TC1, 256-bit key, 12 rounds),
      * rand `src/rngs/std.rs::test_stdrng_construction`: StdRng::from_seed([1,0,0,0, 23,0,0,0, 200,1,0,0, 210,30,0,0, 0...])
        .next_u64() == 10719222850664546238 — which also pins StdRng == ChaCha with 12 rounds.
    """
    import ctypes as C
    import struct

    o = oracle.load()

    def words(key, rounds, n):
        k = (C.c_uint32 * 8)(*key)
        out = (C.c_uint32 * n)()
        o.orc_chacha_words(k, rounds, n, out)
        return list(out)

    z = [0] * 8
    assert struct.pack("<16I", *words(z, 20, 16)).hex() == (
        "76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
        "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    assert struct.pack("<16I", *words(z, 12, 16)).hex() == (
        "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
        "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
    seed = bytes([1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0] + [0] * 16)
    w = words(struct.unpack("<8I", seed), 12, 2)
    assert (w[0] | (w[1] << 32)) == 10719222850664546238


def test_ukeire_golden():
    """shanten.rs:250-393 (calculate_shanten / effective_tiles_with_discard / best_ukeire — what the extended encoders'
    shanten channels call) against answers computed from the reference's own nyanten tables (tests/golden/make_golden.py)."""
    import ctypes as C
    import os

    o = oracle.load()
    path = os.path.join(os.path.dirname(__file__), "golden", "ukeire_golden.txt")
    n = 0
    for line in open(path):
        if line.startswith("#"):
            continue
        hs, vs, es = [x.strip() for x in line.split("|")]
        hand = [int(x) for x in hs.split(",")]
        vis = [int(x) for x in vs.split(",")] if vs else []
        exp = [int(x) for x in es.split()]
        out = (C.c_int * 3)()
        o.orc_ukeire((C.c_int * len(hand))(*hand), len(hand), (C.c_int * max(1, len(vis)))(*vis), len(vis), out)
        assert list(out) == exp, (hand, vis, list(out), exp)
        n += 1
    assert n == 400


def _parse_counts(hs):
    cnt = [0] * 34
    digs = []
    for ch in hs:
        if ch.isdigit():
            digs.append(int(ch))
        else:
            base = {"m": 0, "p": 9, "s": 18, "z": 27}[ch]
            for d in digs:
                cnt[base + d - 1] += 1
            digs = []
    return cnt


def test_shanten_3p_golden_and_known_answers():
    """calculate_shanten_3p (shanten.rs:407-484): 3,000 answers computed from the reference's own tables
    (tests/golden/make_golden.py, incl. the relocation-overflow hands and hands with 2m-8m) and the known answers of the
    reference's tests/test_shanten.py:4-75."""
    import ctypes as C
    import os

    o = oracle.load()

    def sh3(cnt):
        return o.orc_shanten_counts_3p((C.c_uint8 * 34)(*cnt), sum(cnt) // 3)

    def sh4(cnt):
        return o.orc_shanten_counts((C.c_uint8 * 34)(*cnt), sum(cnt) // 3)

    n = 0
    for line in open(os.path.join(os.path.dirname(__file__), "golden", "shanten3p_golden.txt")):
        if line.startswith("#"):
            continue
        digits, exp = line.split()
        cnt = [int(c) for c in digits]
        assert sh3(cnt) == int(exp), (digits, sh3(cnt), exp)
        n += 1
    assert n == 3000
    for hs, e4, e3 in (("1111m111122233z", 1, 2), ("111m111z222z333z44z", -1, -1), ("123456789p11222z", -1, -1),
                       ("111m123456789s11z", -1, -1), ("19m19p19s1234567z", 0, 0), ("111m999m123p789s1z", 0, 0),
                       ("1199m1199p1199s1z", 0, 0), ("11m99m123p456s111z", 0, 0), ("111m999m123p13s7z", 1, 1),
                       ("11119999m22345s", 1, 2), ("1111m9m1234567z", 3, 3), ("111m999m111p11z", -1, -1),
                       ("111m123456789p1z", 0, 0), ("999m111222333z1p", 0, 0), ("11m99m11p99p11s99s1z", 0, 0),
                       ("111999m111999p1z", 0, 0), ("19m147p258s12345z", 5, 5)):
        c = _parse_counts(hs)
        assert sh4(c) == e4 and sh3(c) == e3, hs


def test_ukeire_3p_golden():
    """shanten.rs:470-615 (the 3P shanten / effective-tile / ukeire helpers Observation3P's extended encoders call) against
    answers computed from the reference's own tables (tests/golden/make_golden.py)."""
    import ctypes as C
    import os

    o = oracle.load()
    n = 0
    for line in open(os.path.join(os.path.dirname(__file__), "golden", "ukeire3p_golden.txt")):
        if line.startswith("#"):
            continue
        hs, vs, es = [x.strip() for x in line.split("|")]
        hand = [int(x) for x in hs.split(",")]
        vis = [int(x) for x in vs.split(",")] if vs else []
        out = (C.c_int * 3)()
        o.orc_ukeire_3p((C.c_int * len(hand))(*hand), len(hand), (C.c_int * max(1, len(vis)))(*vis), len(vis), out)
        assert list(out) == [int(x) for x in es.split()], (hand, vis, list(out), es)
        n += 1
    assert n == 300


def test_oracle_3p_extended_encoder_consistency():
    """Observation3P::encode_extended restatement (oracle only so far): over a seeded sanma hanchan the base block equals
    Observation3P::encode except channel 30, the shanten channels equal the pinned 3P helpers, every block keeps its shape
    (fourth relative seat and third opponent empty)."""
    import ctypes as C

    import numpy as np

    from tests.backends import OracleBackend

    o = oracle.load()
    g = OracleBackend(5, 41)
    g.reset()
    ext = np.full(215 * 27 + 8, 7.0, np.float32)
    base = np.zeros(74 * 27, np.float32)
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    n = 0
    while True:
        s = g.get_state()
        if s.is_done:
            break
        for p in range(3):
            if not (s.active_mask >> p) & 1:
                continue
            o.orc_game_encode_ext(g.h, p, fp(ext))
            o.orc_game_encode(g.h, p, fp(base), None)
            assert (ext[215 * 27:] == 7.0).all()
            e, b = ext[:215 * 27].reshape(215, 27), base.reshape(74, 27)
            keep = [c for c in range(74) if c != 30]
            assert (e[keep] == b[keep]).all() and (e[30] >= b[30]).all()
            hand = [s.hand[p][k] for k in range(s.hand_len[p])]
            vis = [s.river[q][k] for q in range(3) for k in range(s.n_river[q])]
            vis += [s.meld_tiles[q][m][k] for q in range(3) for m in range(s.n_melds[q]) for k in range(4) if s.meld_tiles[q][m][k] != 255]
            vis += [s.dora_ind[k] for k in range(s.n_dora)]
            out = (C.c_int * 3)()
            o.orc_ukeire_3p((C.c_int * len(hand))(*hand), len(hand), (C.c_int * max(1, len(vis)))(*vis), len(vis), out)
            f = np.float32
            assert (e[78] == f(max(out[0], 0)) / f(8)).all() and (e[79] == f(out[1]) / f(27)).all() and (e[80] == f(out[2]) / f(80)).all()
            assert not e[77].any() and not e[90:94].any() and not e[97].any() and not e[158:178].any()
            assert not e[203:206].any() and not e[212:215].any()
            assert np.isfinite(e).all()
            n += 1
        g.random_step(3, 41)
    assert n > 300
