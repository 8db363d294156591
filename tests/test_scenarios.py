"""Rule scenarios re-expressed from the reference's own pytest suite (state injection through
snapshot get/set instead of PyO3 setters).  Each scenario runs on the oracle and on the kernel
source compiled for the host; the same scenarios run on the GPU through the C ABI under -m gpu.

Sources (relative to /root/reference/tests): env/rule_validation/test_claim_priority.py,
env/test_illegal_actions.py, test_midway_draw.py, env/test_kan_dora_timing_events.py,
env/actions/test_kyushu_kyuhai.py, env/rule_validation/test_kuikae.py, env/test_riichienv.py.
"""
import json

import pytest

from riichienv_b200 import _abi as A
from tests.backends import BACKENDS, act, setup_env

CPU_BACKENDS = ["oracle", "hostsim"]
ALL = [pytest.param(b) for b in CPU_BACKENDS] + [pytest.param("gpu", marks=pytest.mark.gpu)]


def ev(env):
    return [json.loads(s) for s in env.events_json()]


def active(s):
    return [p for p in range(4) if (s.active_mask >> p) & 1]


@pytest.mark.parametrize("backend", ALL)
def test_initialization(backend):  # env/test_riichienv.py:9-64
    env = BACKENDS[backend](0, 42)
    env.reset()
    s = env.get_state()
    assert s.wall_top - s.rinshan_draw_count == 83
    assert [s.hand_len[p] for p in range(4)] == [14, 13, 13, 13]
    assert s.current_player == 0 and s.turn_count == 0 and not s.is_done and not s.needs_tsumo
    assert s.drawable_count == 69
    assert active(s) == [0]
    assert len(env.legal_tuples(0)) == 14 and env.legal_tuples(1) == []
    e = ev(env)
    assert [x["type"] for x in e] == ["start_game", "start_kyoku", "tsumo"]
    masked = [json.loads(x) for x in env.events_json(viewer=1)]
    assert masked[1]["tehais"][0][0] == "?" and masked[1]["tehais"][1][0] != "?" and masked[2]["pai"] == "?"


@pytest.mark.parametrize("backend", ALL)
def test_basic_step(backend):  # env/test_riichienv.py:66-126
    env = BACKENDS[backend](0, 42)
    env.reset()
    s = env.get_state()
    tile = s.hand[0][s.hand_len[0] - 1]
    env.step({0: act(A.DISCARD, tile)})
    while env.get_state().phase == 1:
        env.step({p: act(A.PASS) for p in active(env.get_state())})
    s = env.get_state()
    assert s.phase == 0 and s.current_player == 1 and s.hand_len[1] == 14 and s.drawn_tile != 255
    assert [x["type"] for x in ev(env)][:5] == ["start_game", "start_kyoku", "tsumo", "dahai", "tsumo"]


@pytest.mark.parametrize("backend", ALL)
def test_pon_priority_over_chi(backend):  # env/rule_validation/test_claim_priority.py:11-54
    env = setup_env(BACKENDS[backend], seed=1,
                    hands=[[57] + [2] * 12, [62, 65] + [0] * 11, [56, 58] + [1] * 11,
                           [12, 16, 19, 21, 48, 59, 64, 77, 81, 89, 104, 130, 133]],
                    current_player=0, active_players=[0], drawn_tile=100)
    env.step({0: act(A.DISCARD, 57)})
    s = env.get_state()
    assert s.phase == 1 and active(s) == [1, 2]
    env.step({1: act(A.CHI, 57, [62, 65]), 2: act(A.PON, 57, [56, 58])})
    s = env.get_state()
    assert s.phase == 0 and active(s) == [2]
    assert ev(env)[-1]["type"] == "pon"


@pytest.mark.parametrize("backend", ALL)
def test_illegal_discard_penalty(backend):  # env/test_illegal_actions.py:5-60
    env = BACKENDS[backend](1, 42)
    env.reset()
    s = env.get_state()
    hand = [s.hand[0][k] for k in range(s.hand_len[0])]
    bad = 0
    while bad in hand:
        bad += 1
    env.step({0: act(A.DISCARD, bad)})
    s = env.get_state()
    assert s.last_error == 0 and not s.is_done
    e = ev(env)
    ry = [x for x in e if x["type"] == "ryukyoku"][-1]
    assert "Error: Illegal Action" in ry["reason"] and ry["deltas"] == [-12000, 4000, 4000, 4000]
    assert [s.score[p] for p in range(4)] == [13000, 29000, 29000, 29000]
    assert s.oya == 0 and s.honba == 1 and s.kyoku_idx == 0 and s.hand_len[0] == 14
    assert any(x["type"] == "end_kyoku" for x in e)


@pytest.mark.parametrize("backend", ALL)
def test_illegal_out_of_turn(backend):  # env/test_illegal_actions.py:62-100 (ko offender: -8000 / oya +4000 / ko +2000)
    env = BACKENDS[backend](1, 42)
    env.reset()
    s = env.get_state()
    env.step({0: act(A.DISCARD, s.hand[0][13]), 1: act(A.DISCARD, 0)})
    s = env.get_state()
    assert s.last_error == 1
    assert [s.score[p] for p in range(4)] == [29000, 17000, 27000, 27000]


@pytest.mark.parametrize("backend", ALL)
def test_sufuurenta(backend):  # test_midway_draw.py:7-31
    tiles = [108, 109, 110, 111]
    scattered = [0, 4, 8, 36, 40, 44, 72, 76, 80, 112, 116, 120, 124]
    env = setup_env(BACKENDS[backend], hands=[scattered[:] for _ in range(4)], wall=list(range(136)))
    for i in range(4):
        s = env.get_state()
        p = s.current_player
        hand = sorted([s.hand[p][k] for k in range(s.hand_len[p])])
        hand[0] = tiles[i]
        for k in range(s.hand_len[p]):
            s.hand[p][k] = hand[k]
        s.drawn_tile = tiles[i]
        env.set_state(s)
        env.step({p: act(A.DISCARD, tiles[i])})
        if i < 3:
            assert not env.get_state().is_done
    assert env.get_state().is_done
    assert any(x.get("reason") == "sufuurenta" for x in ev(env))


@pytest.mark.parametrize("backend", ALL)
def test_suukansansen(backend):  # test_midway_draw.py:33-56
    scattered = [0, 4, 8, 36, 40, 44, 72, 76, 80, 112, 116, 120, 124]
    ank = lambda lo: (3, [lo, lo + 1, lo + 2, lo + 3], -1, None)
    env = setup_env(BACKENDS[backend], hands=[scattered[:] for _ in range(4)],
                    melds=[[ank(0), ank(4)], [ank(8), ank(12)], [], []], current_player=1, drawn_tile=108,
                    wall=list(range(136)))
    env.step({1: act(A.DISCARD, 108)})
    assert env.get_state().is_done
    assert any(x.get("reason") == "suukansansen" for x in ev(env))


@pytest.mark.parametrize("backend", ALL)
def test_kyushu_kyuhai(backend):  # env/actions/test_kyushu_kyuhai.py
    hand = [0, 32, 36, 68, 72, 104, 108, 112, 116, 4, 8, 12, 40]  # 9 distinct terminal/honor kinds
    env = setup_env(BACKENDS[backend], hands=[hand, None, None, None], current_player=0, drawn_tile=44)
    legal = env.legal_tuples(0)
    assert (A.KYUSHU_KYUHAI, None, ()) in legal
    env.step({0: act(A.KYUSHU_KYUHAI)})
    e = ev(env)
    assert any(x.get("reason") == "kyushu_kyuhai" for x in e)
    assert env.get_state().is_done  # single-kyoku mode


@pytest.mark.parametrize("backend", ALL)
def test_kuikae_forbidden_after_chi(backend):  # env/rule_validation/test_kuikae.py
    # P1 holds 2m3m (+ 1m,4m); P0 discards 4m -> chi 2m3m+4m forbids discarding 4m and 1m (suji kuikae)
    p1 = [0 * 4, 1 * 4, 2 * 4, 3 * 4 + 1, 40, 44, 72, 76, 80, 112, 116, 120, 124]
    p0 = [3 * 4, 41, 45, 73, 77, 81, 113, 117, 121, 125, 100, 101, 102]
    env = setup_env(BACKENDS[backend], hands=[p0, p1, None, None], current_player=0, drawn_tile=133)
    env.step({0: act(A.DISCARD, 12)})
    s = env.get_state()
    assert s.phase == 1 and 1 in active(s)
    chis = [a for a in env.legal_tuples(1) if a[0] == A.CHI]
    assert (A.CHI, 12, (4, 8)) in chis
    acts = {p: act(A.PASS) for p in active(s)}
    acts[1] = act(A.CHI, 12, [4, 8])
    env.step(acts)
    s = env.get_state()
    assert s.phase == 0 and s.current_player == 1
    discards = [a[1] for a in env.legal_tuples(1) if a[0] == A.DISCARD]
    assert 13 not in discards and 0 not in discards  # 4m (other copy) and 1m are forbidden
    assert 40 in discards


@pytest.mark.parametrize("backend", ALL)
def test_kan_dora_timing(backend):  # env/test_kan_dora_timing_events.py
    # ankan: dora event BEFORE the rinshan tsumo
    h0 = [0, 1, 2, 36, 40, 44, 72, 76, 80, 112, 116, 120, 124]
    env = setup_env(BACKENDS[backend], hands=[h0, None, None, None], current_player=0, drawn_tile=3, wall=list(range(136)))
    assert (A.ANKAN, 0, (0, 1, 2, 3)) in env.legal_tuples(0)
    env.step({0: act(A.ANKAN, 0, [0, 1, 2, 3])})
    types = [x["type"] for x in ev(env)]
    assert types[-3:] == ["ankan", "dora", "tsumo"]
    s = env.get_state()
    assert s.n_dora == 2 and s.rinshan_draw_count == 1 and s.is_rinshan_flag == 1
    # kakan: tsumo first, dora revealed right before the next dahai
    pon = (1, [4, 5, 6], 1, 4)
    h0 = [7, 36, 40, 44, 72, 76, 80, 112, 116, 120]
    env = setup_env(BACKENDS[backend], hands=[h0, None, None, None], melds=[[pon], [], [], []], current_player=0,
                    drawn_tile=124, wall=list(range(136)))
    kakans = [a for a in env.legal_tuples(0) if a[0] == A.KAKAN]
    assert kakans == [(A.KAKAN, 7, (4, 5, 6))]
    env.step({0: act(A.KAKAN, 7, [4, 5, 6])})
    s = env.get_state()
    if s.phase == 1:  # somebody may chankan: everyone passes
        env.step({p: act(A.PASS) for p in active(s)})
    types = [x["type"] for x in ev(env)]
    assert types[-2:] == ["kakan", "tsumo"]
    s = env.get_state()
    assert s.pending_kan_dora_count == 1 and s.n_dora == 1
    env.step({0: act(A.DISCARD, s.drawn_tile)})
    types = [x["type"] for x in ev(env)]
    i = max(k for k, t in enumerate(types) if t == "dahai")
    assert types[i - 1] == "dora"


@pytest.mark.parametrize("backend", ALL)
def test_ron_and_scores(backend):  # tests/test_env_scoring.py style: P1 rons P0's discard
    # P1: 123m 456m 789m 11p + 23p waiting 1p/4p ; menzen pinfu-ish + ittsu
    p1 = [0, 4, 8, 12, 17, 20, 24, 28, 32, 36, 37, 40, 44]
    p0 = [48, 72, 76, 80, 84, 88, 112, 116, 120, 124, 128, 132, 100]
    env = setup_env(BACKENDS[backend], hands=[p0, p1, None, None], current_player=0, drawn_tile=52, game_mode=1)
    env.step({0: act(A.DISCARD, 48)})  # 4p
    s = env.get_state()
    assert s.phase == 1 and 1 in active(s)
    assert (A.RON, 48, ()) in env.legal_tuples(1)
    env.step({p: (act(A.RON, 48) if p == 1 else act(A.PASS)) for p in active(s)})
    e = ev(env)
    hora = [x for x in e if x["type"] == "hora"][-1]
    assert hora["actor"] == 1 and hora["target"] == 0 and sum(hora["deltas"]) == 0 and hora["deltas"][1] > 0
    s = env.get_state()
    assert s.score[0] + s.score[1] == 50000 and s.oya == 1 and s.honba == 0


@pytest.mark.parametrize("backend", ALL)
def test_tsumo_and_riichi_flow(backend):  # env/rule_validation/test_riichi_sequence.py style
    # P0 tenpai after discarding 9s: 123m 456m 789m 11p 23p + junk 9s; riichi -> discard -> accepted
    p0 = [0, 4, 8, 12, 17, 20, 24, 28, 32, 36, 37, 40, 44]
    env = setup_env(BACKENDS[backend], hands=[p0, None, None, None], current_player=0, drawn_tile=104, wall=list(range(136)))
    legal = env.legal_tuples(0)
    assert (A.RIICHI, None, ()) in legal
    env.step({0: act(A.RIICHI)})
    s = env.get_state()
    assert s.flags[0] & A.F_RIICHI_STAGE
    legal = env.legal_tuples(0)
    # tenpai-keeping discards only: 9s (wait 1p/4p) or a 1p (tanki on 9s)
    assert [a for a in legal if a[0] == A.DISCARD] == [(A.DISCARD, 36, ()), (A.DISCARD, 37, ()), (A.DISCARD, 104, ())]
    assert A.RIICHI not in [a[0] for a in legal] and A.TSUMO not in [a[0] for a in legal]
    env.step({0: act(A.DISCARD, 104)})
    s = env.get_state()
    while s.phase == 1:
        env.step({p: act(A.PASS) for p in active(s)})
        s = env.get_state()
    types = [x["type"] for x in ev(env)]
    assert "reach" in types and "reach_accepted" in types
    assert types.index("reach") < types.index("reach_accepted")
    assert s.score[0] == 24000 and s.riichi_sticks == 1 and (s.flags[0] & A.F_RIICHI_DECLARED)
    assert s.flags[0] & A.F_DOUBLE_RIICHI  # declared on the first turn


# ---------------------------------------------------------------------------------------------------------
# 3-player (sanma) scenarios — re-expressed from tests/env/test_sanma.py


@pytest.mark.parametrize("backend", ALL)
def test_sanma_initialization(backend):  # test_sanma.py:50-118
    for mode in (3, 4, 5):
        env = BACKENDS[backend](mode, 42)
        env.reset()
        s = env.get_state()
        assert [s.hand_len[p] for p in range(4)] == [14, 13, 13, 0]
        assert [s.score[p] for p in range(3)] == [35000, 35000, 35000]
        assert s.wall_top - s.rinshan_draw_count == 68 and s.wall_len == 108 and s.drawable_count == 54
        tiles = [s.hand[p][k] for p in range(3) for k in range(s.hand_len[p])] + [s.wall[i] for i in range(s.wall_top)]
        assert all(not (1 <= t // 4 <= 7) for t in tiles) and len(set(tiles)) == 108   # no 2m-8m
        e = ev(env)
        assert [x["type"] for x in e] == ["start_game", "start_kyoku", "tsumo"]
        assert len(e[1]["tehais"]) == 3 and len(e[1]["scores"]) == 3


@pytest.mark.parametrize("backend", ALL)
def test_sanma_rotation_and_no_chi(backend):  # test_sanma.py:120-175
    env = BACKENDS[backend](5, 42)
    env.reset()
    for expected in (0, 1, 2, 0):
        s = env.get_state()
        assert s.current_player == expected
        assert all(a[0] != A.CHI for a in env.legal_tuples(expected))
        env.step({expected: act(A.DISCARD, s.hand[expected][s.hand_len[expected] - 1])})
        while env.get_state().phase == 1:
            st = env.get_state()
            for p in active(st):
                assert all(a[0] != A.CHI for a in env.legal_tuples(p))
            env.step({p: act(A.PASS) for p in active(st)})
    # shimocha holding 2p3p cannot chi a 1p
    h0 = [36, 40, 44, 48, 52, 56, 60, 64, 68, 72, 76, 80, 84]
    h1 = [37, 41, 45, 49, 53, 57, 61, 65, 69, 73, 77, 81, 85]
    env = setup_env(BACKENDS[backend], game_mode=5, hands=[h0, h1, None], current_player=0, drawn_tile=88)
    env.step({0: act(A.DISCARD, 36)})
    st = env.get_state()
    for p in active(st) if st.phase == 1 else []:
        assert all(a[0] != A.CHI for a in env.legal_tuples(p))


@pytest.mark.parametrize("backend", ALL)
def test_sanma_pon_claim(backend):  # test_sanma.py:177-205
    h0 = [36, 40, 44, 48, 52, 56, 60, 64, 68, 72, 76, 80, 84]
    h1 = [37, 38, 49, 53, 57, 61, 65, 69, 73, 77, 81, 85, 89]
    env = setup_env(BACKENDS[backend], game_mode=5, hands=[h0, h1, None], current_player=0, drawn_tile=88)
    env.step({0: act(A.DISCARD, 36)})
    s = env.get_state()
    assert s.phase == 1 and 1 in active(s)
    pons = [a for a in env.legal_tuples(1) if a[0] == A.PON]
    assert pons == [(A.PON, 36, (37, 38))]
    acts = {p: act(A.PASS) for p in active(s)}
    acts[1] = act(A.PON, 36, [37, 38])
    env.step(acts)
    s = env.get_state()
    assert s.current_player == 1 and s.phase == 0 and s.n_melds[1] == 1


@pytest.mark.parametrize("backend", ALL)
def test_sanma_tsumo_and_ron_deltas(backend):  # test_sanma.py:403-466
    hand = [36, 37, 38, 40, 41, 42, 44, 45, 46, 32, 33, 34, 0]
    env = setup_env(BACKENDS[backend], game_mode=5, hands=[hand, None, None], current_player=0, drawn_tile=1,
                    discards=[[100], [], []])
    s = env.get_state()
    s.is_first_turn = 0
    env.set_state(s)
    assert (A.TSUMO, 1, ()) in env.legal_tuples(0)
    env.step({0: act(A.TSUMO)})
    hora = [x for x in ev(env) if x["type"] == "hora"][-1]
    d = hora["deltas"]
    assert hora["tsumo"] is True and len(d) == 3 and d[0] > 0 and d[1] < 0 and d[2] < 0 and sum(d) == 0
    # ron
    p0 = [1, 48, 52, 56, 60, 64, 68, 72, 76, 80, 84, 88, 92]
    env = setup_env(BACKENDS[backend], game_mode=5, hands=[p0, hand, None], current_player=0, drawn_tile=96)
    env.step({0: act(A.DISCARD, 1)})
    s = env.get_state()
    assert s.phase == 1 and 1 in active(s)
    assert [a for a in env.legal_tuples(1) if a[0] == A.RON] == [(A.RON, 1, ())]
    env.step({p: (act(A.RON, 1) if p == 1 else act(A.PASS)) for p in active(s)})
    d = [x for x in ev(env) if x["type"] == "hora"][-1]["deltas"]
    assert len(d) == 3 and d[1] > 0 and d[0] < 0 and d[2] == 0 and sum(d) == 0


@pytest.mark.parametrize("backend", ALL)
def test_sanma_kita(backend):  # state_3p/sanma.rs:9-204, tests/env/test_sanma.py (kita flow)
    hand = [36, 40, 44, 48, 52, 56, 60, 64, 68, 72, 76, 80, 120]   # holds N (120)
    env = setup_env(BACKENDS[backend], game_mode=5, hands=[hand, None, None], current_player=0, drawn_tile=121)
    legal = env.legal_tuples(0)
    kitas = [a for a in legal if a[0] == A.KITA]
    assert kitas == [(A.KITA, 120, ()), (A.KITA, 121, ())] and legal[-1][0] == A.KITA
    before = env.get_state()
    env.step({0: act(A.KITA, 120)})
    s = env.get_state()
    if s.phase == 1:   # somebody may ron the kita tile: everyone passes
        env.step({p: act(A.PASS) for p in active(s)})
        s = env.get_state()
    assert s.n_kita[0] == 1 and s.hand_len[0] == 14 and s.rinshan_draw_count == 1 and s.is_rinshan_flag == 1
    assert s.drawable_count == before.drawable_count - 1 and s.n_dora == 1      # no new dora for kita
    types = [x["type"] for x in ev(env)]
    assert types[-2:] == ["kita", "tsumo"]
    assert s.is_first_turn == 0


# ---- extended encoders: the reference's own unit tests (observation/encode.rs:593-800) -----------------------------------
_OBS_HANDS = [[0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48], [1, 5, 9, 13, 17, 21, 25, 29, 33, 37, 41, 45, 49],
              [2, 6, 10, 14, 18, 22, 26, 30, 34, 38, 42, 46, 50], [3, 7, 11, 15, 19, 23, 27, 31, 35, 39, 43, 47, 51]]


def _ext(backend, pid, discards=None, melds=None):
    """make_obs (encode.rs:593-621) as a live state: seat `pid` is the one observed (it owes the action)."""
    env = setup_env(BACKENDS[backend], seed=3, hands=_OBS_HANDS, current_player=pid, active_players=[pid],
                    discards=discards or [[], [], [], []], melds=melds)
    return env.encode_ext(pid)


@pytest.mark.parametrize("backend", ALL)
def test_ext_discard_decay_relative_order(backend):  # encode.rs:633-662, 785-803
    d = [[0], [], [36], []]
    b0, b2 = _ext(backend, 0, d), _ext(backend, 2, d)
    assert b0[74, 0] > 0 and b0[76, 9] > 0
    assert b2[74, 9] > 0 and b2[76, 0] > 0
    assert b0[74, 0] == b2[74, 9] == 1.0
    d = [[0], [4], [8], [12]]
    for pid in range(4):
        assert _ext(backend, pid, d)[74, d[pid][0] // 4] > 0
    # exp(-0.2 * age): two discards of one kind accumulate, oldest first
    import numpy as np
    b = _ext(backend, 1, [[], [0, 1, 40], [], []])
    w = [np.float32(np.exp(np.float32(-0.2) * np.float32(a))) for a in (2, 1)]
    assert abs(b[74, 0] - (w[0] + w[1])) < 1e-6 and b[74, 10] == 1.0


@pytest.mark.parametrize("backend", ALL)
def test_ext_shanten_relative_order(backend):  # encode.rs:665-702
    d = [[0, 4], [8], [12, 16, 20], []]
    b0, b2 = _ext(backend, 0, d), _ext(backend, 2, d)
    assert abs(b0[78 + 3, 0] - b2[78 + 2 * 4 + 3, 0]) < 1e-6
    assert b0[78 + 4, 0] == 0.5 and abs(b0[78, 0] - 0.5) > 1e-6
    assert b2[78 + 4, 0] == 0.5 and abs(b2[78, 0] - 0.5) > 1e-6
    assert (b0[78 + 3] == b0[78 + 3, 0]).all()      # broadcast


@pytest.mark.parametrize("backend", ALL)
def test_ext_ankan_and_fuuro_relative_order(backend):  # encode.rs:705-782
    ankan = [None, [(3, [0, 1, 2, 3], 0, None)], None, None]
    b0, b3 = _ext(backend, 0, melds=ankan), _ext(backend, 3, melds=ankan)
    assert b0[94 + 1, 0] == 1.0 and b0[94, 0] == 0.0
    assert b3[94 + 2, 0] == 1.0 and b3[94 + 1, 0] == 0.0
    chi = [None, None, [(0, [0, 4, 8], 1, 0)], None]
    b0, b1 = _ext(backend, 0, melds=chi), _ext(backend, 1, melds=chi)
    assert b0[98 + 40, 0] == 1.0 and b0[98 + 41, 1] == 1.0
    assert b1[98 + 20, 0] == 1.0 and b1[98 + 21, 1] == 1.0
    # a called meld counts one tile short in channel 30 of encode_base_into (encode.rs:94-110)
    import numpy as np
    used = 13 + 1 + 3 - 1      # own hand, the dora indicator, the chi minus its called tile
    assert b0[30, 0] == np.float32(136 - used) / np.float32(70.0)


@pytest.mark.parametrize("backend", ALL)
def test_ext_tile_context_channels(backend):
    """pass context / last tedashi / riichi sutehai (encode.rs:479-584) from injected state: opponents in ABSOLUTE seat order,
    "is dora" only for copy 0 of the dora kind (a tile id is compared with get_next_tile's id), and the pass context is fed
    the discarder's SEAT (state/mod.rs:252 destructures the (pid, tile) tuple the other way round)."""
    import numpy as np

    env = setup_env(BACKENDS[backend], seed=3, hands=_OBS_HANDS, current_player=1, active_players=[1])
    s = env.get_state()
    s.n_dora = 1
    s.dora_ind[0] = 4 * 12 + 2          # indicator 4p -> dora 5p (kind 13): only tile id 52 "is dora"
    s.riichi_sutehai[0], s.riichi_sutehai[2], s.riichi_sutehai[3] = 52, 53, 255
    s.last_tedashi[0], s.last_tedashi[2], s.last_tedashi[3] = 255, 88, 135
    s.last_discard_pid, s.last_discard_tile = 2, 52
    env.set_state(s)
    b = env.encode_ext(1)
    f = np.float32
    # riichi sutehai: opponents of seat 1 in absolute order are 0, 2, 3
    assert (b[206] == f(13) / f(33)).all() and (b[207] == 1).all() and (b[208] == 1).all()     # 52: red 5p, copy 0 -> dora
    assert (b[209] == f(13) / f(33)).all() and (b[210] == 0).all() and (b[211] == 0).all()     # 53: same kind, not copy 0
    assert not b[212:215].any()
    # last tedashi
    assert not b[197:200].any()
    assert (b[200] == f(22) / f(33)).all() and (b[201] == 1).all() and (b[202] == 0).all()     # 88: red 5s
    assert (b[203] == f(33) / f(33)).all() and (b[204] == 0).all() and (b[205] == 0).all()
    # pass context sees "tile" 2 (the discarder's seat): kind 0, not red, not dora
    assert not b[194:197].any()
    s.dora_ind[0] = 4 * 8               # indicator 9m -> dora 1m (kind 0, copy 0 = tile id 0)
    s.last_discard_pid = 0
    env.set_state(s)
    b = env.encode_ext(1)
    assert (b[194] == 0).all() and (b[195] == 0).all() and (b[196] == 1).all()                  # seat 0 "is" tile id 0


# ---- round / game flow (state/mod.rs:1595-1688), re-expressed from the reference's Rust unit tests and tests/test_oyayame_tiebreak.py.
# The Rust tests call _trigger_ryukyoku / _initialize_next_round directly on a GameState; ops 2-5 of the debug call do the same here.
def _flow_env(backend, mode=2, **fields):
    env = BACKENDS[backend](mode, 7)
    env.reset()
    s = env.get_state()
    for k, v in fields.items():
        if k == "scores":
            for p, x in enumerate(v):
                s.score[p] = x
        elif k == "nagashi_off":
            for p in range(4):
                s.flags[p] &= ~A.F_NAGASHI_ELIGIBLE
        else:
            setattr(s, k, v)
    env.set_state(s)
    return env


@pytest.mark.parametrize("backend", ALL)
def test_sudden_death_hanchan_logic(backend):  # riichienv-core/src/tests.rs:172-233
    env = _flow_env(backend, round_wind=1, kyoku_idx=3, oya=3, scores=[25000] * 4, nagashi_off=True)
    assert env.call(2) == 0                       # exhaustive draw in South 4 with nobody at 30000: the game goes on
    s = env.get_state()
    assert (s.round_wind, s.kyoku_idx, s.oya, s.is_done) == (2, 0, 0, 0)     # West 1, dealer seat 0
    for p, x in enumerate([31000, 25000, 24000, 20000]):
        s.score[p] = x
    for p in range(4):
        s.flags[p] &= ~A.F_NAGASHI_ELIGIBLE
    env.set_state(s)
    assert env.call(2) == 1                       # somebody at 30000 in the West round: over
    types = [x["type"] for x in ev(env)]
    assert types[-1] == "end_game" and "ryukyoku" in types


@pytest.mark.parametrize("backend", ALL)
def test_tobi_ends_game(backend):  # tests.rs:375-409
    env = _flow_env(backend, scores=[30000, 40000, 35000, -5000])
    assert env.call(3) == 1
    types = [x["type"] for x in ev(env)]
    assert "end_kyoku" in types and types.index("end_kyoku") < types.index("end_game")


@pytest.mark.parametrize("backend", ALL)
def test_oyayame_requires_target_in_orasu(backend):  # tests.rs:411-427, tests/test_oyayame_tiebreak.py:75-87
    env = _flow_env(backend, round_wind=1, oya=3, kyoku_idx=3, scores=[28900, 20000, 20100, 29000])
    assert env.call(4) == 0                       # dealer top but below 30000: renchan, the game continues
    s = env.get_state()
    assert (s.oya, s.round_wind, s.honba) == (3, 1, 1)


@pytest.mark.parametrize("backend", ALL)
def test_oyayame_tiebreak_last_round(backend):  # tests/test_oyayame_tiebreak.py:44-72 on the env's own rule
    # South 4, dealer (seat 3) tied with seat 0 at 30000: the tie goes to the lower seat, the dealer is not top -> no agari-yame
    env = _flow_env(backend, round_wind=1, oya=3, kyoku_idx=3, scores=[30000, 20000, 20000, 30000])
    assert env.call(4) == 0
    # dealer sole top at 30000 -> the game ends on the dealer's win
    env = _flow_env(backend, round_wind=1, oya=3, kyoku_idx=3, scores=[29000, 20000, 21000, 30000])
    assert env.call(4) == 1
    # dealer seat 0 ties seat 1 for top (ties go to the lower seat = the dealer): only the LAST dealer can stop; South 1 goes on
    env = _flow_env(backend, round_wind=1, oya=0, kyoku_idx=0, scores=[35000, 35000, 15000, 15000])
    assert env.call(4) == 0


def test_yaku_possibility_known_answers():
    """riichienv-core/src/yaku_checker.rs:413-467 (the reference's unit tests) and the shape of encode_yaku_possibility"""
    from types import SimpleNamespace as NS

    import numpy as np

    from riichienv_b200 import yaku_possibility as Y

    M = lambda tiles: NS(tiles=list(tiles))
    assert Y.check_tanyao([]) == Y.UNKNOWN
    assert Y.check_tanyao([M([0, 1, 2])]) == Y.IMPOSSIBLE          # 1m pon
    assert Y.check_tanyao([M([16, 17, 18])]) == Y.UNKNOWN           # 5m pon
    assert Y.check_toitoi([M([0, 4, 8])]) == Y.IMPOSSIBLE           # 1m-2m-3m
    assert Y.check_toitoi([M([16, 17, 18])]) == Y.POSSIBLE
    assert Y.check_yakuhai(31, [M([124, 125, 126])], [], []) == Y.POSSIBLE
    assert Y.check_yakuhai(31, [], [124, 125], [126]) == Y.IMPOSSIBLE      # three of four visible
    assert Y.check_flush([M([0, 4, 8]), M([108, 109, 110])]) == (Y.POSSIBLE, Y.IMPOSSIBLE)
    assert Y.check_flush([M([0, 4, 8]), M([36, 40, 44])]) == (Y.IMPOSSIBLE, Y.IMPOSSIBLE)
    assert Y.check_daisangen([], [124, 125], []) == Y.IMPOSSIBLE and Y.check_kokushi([M([0, 1, 2])], [], []) == Y.IMPOSSIBLE
    obs = NS(melds=[[M([0, 4, 8])], [], [], []], discards=[[], [124, 125, 126], [], []], dora_indicators=[0], round_wind=0, oya=1)
    a = np.frombuffer(Y.encode(obs, 4), np.float32).reshape(4, 21, 2)
    assert a.shape == (4, 21, 2) and (a[:, :, 0] == a[:, :, 1]).all()
    assert a[0, 0, 0] == 0.0 and a[0, 8, 0] == 0.0 and a[0, 9, 0] == 0.0 and a[0, 19, 0] == 0.0       # tanyao, toitoi, chiitoi, iipeikou
    assert a[1, 1, 0] == 0.0 and a[0, 1, 0] == 1.0 and a[2].min() == 1.0                             # seat 1 discarded three haku
    assert len(Y.encode(NS(melds=[[], [], []], discards=[[], [], []], dora_indicators=[], round_wind=0, oya=0), 3)) == 3 * 21 * 2 * 4
