"""Kernel-source-vs-oracle parity on CPU (no GPU needed): the CUDA device code of
riichienv_b200/csrc is compiled for the host by tests/hostsim and diffed against the oracle.
The same comparisons run against the real GPU library in tests/test_gpu_parity.py."""
import ctypes as C

import pytest

import oracle
from riichienv_b200 import _abi as A
from tests import helpers as H
from tests import hostsim
from tests.backends import HostsimBackend, OracleBackend


@pytest.fixture(scope="module")
def libs():
    return oracle.load(), hostsim.load()


def test_hand_eval_golden(libs):
    orc, hs = libs
    for i, (q, exp, yaku) in enumerate(H.load_agari_cases()):
        r, ro = A.HandResult(), A.HandResult()
        hs.hs_hand_eval(C.byref(q), C.byref(r), 1)
        orc.orc_hand_eval(C.byref(q), C.byref(ro), 1)
        assert (r.is_win, r.han, r.fu) == exp and H.yaku_ids(r.yaku_mask) == yaku, f"case {i}"
        assert bytes(r) == bytes(ro), f"case {i}"


def test_hand_eval_golden_3p(libs):
    import os

    orc, hs = libs
    cases = H.load_agari_cases("agari_3p.txt")
    lines = [l for l in open(os.path.join(H.GOLDEN, "agari_3p.txt")) if not l.startswith("#")]
    for i, ((q, exp, yaku), line) in enumerate(zip(cases, lines)):
        q.sanma, q.kita_count = 1, [int(x) for x in line.split("|")[5].split()][4]
        r, ro = A.HandResult(), A.HandResult()
        hs.hs_hand_eval(C.byref(q), C.byref(r), 1)
        orc.orc_hand_eval(C.byref(q), C.byref(ro), 1)
        assert (r.is_win, r.han, r.fu) == exp and H.yaku_ids(r.yaku_mask) == yaku, f"case {i}"
        assert bytes(r) == bytes(ro), f"case {i}"


def test_hand_eval_random(libs):
    orc, hs = libs
    qs = H.random_hand_queries(20000, seed=7)
    arr = H.query_array(qs)
    a = (A.HandResult * len(qs))()
    b = (A.HandResult * len(qs))()
    hs.hs_hand_eval(arr, a, len(qs))
    orc.orc_hand_eval_mt(arr, b, len(qs), 8)
    wins = 0
    for i in range(len(qs)):
        assert bytes(a[i]) == bytes(b[i]), f"hand {i}"
        wins += a[i].is_win
    assert wins > 1000  # the positive stratum is exercised


def test_shanten_and_waits_golden(libs):
    _, hs = libs
    for cnt, sh in H.load_counts_file("shanten_golden.txt"):
        assert hs.hs_shanten_counts((C.c_uint8 * 34)(*cnt), sum(cnt) // 3) == sh
    for cnt, tenpai in H.load_counts_file("hands_negative.txt"):
        arr = (C.c_uint8 * 34)(*cnt)
        if sum(cnt) == 14:
            assert hs.hs_is_agari(arr) == 0
        else:
            assert (hs.hs_waits(arr) != 0) == bool(tenpai)


def test_wall_matches_oracle(libs):
    orc, hs = libs
    for seed in (0, 1, 2, 42, 123, 2 ** 32, 2 ** 63 + 1):
        for hi in (0, 1, 5):
            for n in (136, 108):
                a = (C.c_uint8 * 136)()
                b = (C.c_uint8 * 136)()
                orc.orc_wall_from_seed(seed, hi, n, a)
                hs.hs_wall_from_seed(seed, hi, n, b)
                assert bytes(a)[:n] == bytes(b)[:n]
                assert len(set(bytes(a)[:n])) == n


def lockstep(seed, mode, rule, agent_seed, check_legal=True, policy=0):
    o, h = OracleBackend(mode, seed, rule), HostsimBackend(mode, seed, rule)
    o.reset()
    h.reset()
    steps = 0
    while True:
        so, sh = o.get_state(), h.get_state()
        d = A.state_fields_equal(so, sh)
        assert not d, f"seed {seed} step {steps}: state differs in {d}"
        if so.is_done:
            break
        if check_legal:
            for p in range(4):
                assert o.legal_tuples(p) == h.legal_tuples(p), f"seed {seed} step {steps} seat {p}"
        if policy:
            o.agent_step(policy, agent_seed, seed)
            h.agent_step(policy, agent_seed, seed)
        else:
            o.random_step(agent_seed, seed)
            h.random_step(agent_seed, seed)
        steps += 1
    assert o.events() == h.events()
    for viewer in (-1, 0, 1, 2):
        assert o.events_json(viewer) == h.events_json(viewer)   # the oracle's own text log vs the product renderer
    return steps


@pytest.mark.parametrize("mode,rule,n", [(2, A.RULE_DEFAULT_TENHOU, 24), (2, A.RULE_DEFAULT_MJSOUL, 12),
                                          (1, A.RULE_DEFAULT_TENHOU, 8), (0, A.RULE_DEFAULT_TENHOU, 16),
                                          (5, A.RULE_DEFAULT_TENHOU, 16), (5, A.RULE_DEFAULT_MJSOUL, 8),
                                          (4, A.RULE_DEFAULT_TENHOU, 8), (3, A.RULE_DEFAULT_TENHOU, 16)])
def test_random_games_lockstep(mode, rule, n):
    total = 0
    for seed in range(100 * mode, 100 * mode + n):
        total += lockstep(seed, mode, rule, agent_seed=0xC0FFEE + mode)
    assert total > 40 * n


@pytest.mark.parametrize("mode,rule,n", [(2, A.RULE_DEFAULT_TENHOU, 16), (2, A.RULE_DEFAULT_MJSOUL, 16), (5, A.RULE_DEFAULT_TENHOU, 12),
                                          (5, A.RULE_DEFAULT_MJSOUL, 8), (0, A.RULE_DEFAULT_MJSOUL, 24), (3, A.RULE_DEFAULT_TENHOU, 24)])
def test_greedy_agent_games_lockstep(mode, rule, n):
    """The greedy-win agent (every Tsumo / Ron / Riichi, calls with probability 1/4, discards towards the lowest shanten):
    ~60 % of the rounds end in a win, so this is the lock-step gate of the settlement path (state/mod.rs:685-893, 919-1142).
    Full record after every step, every legal list, the event words and the MJAI text of the kernel code vs the oracle."""
    total = wins = 0
    for seed in range(500 + 100 * mode, 500 + 100 * mode + n):
        total += lockstep(seed, mode, rule, agent_seed=0xBEEF + mode, policy=1)
    assert total > 40 * n


@pytest.mark.parametrize("mode,rule,n", [(2, A.RULE_DEFAULT_TENHOU, 12), (5, A.RULE_DEFAULT_TENHOU, 12), (1, A.RULE_DEFAULT_MJSOUL, 6),
                                          (4, A.RULE_DEFAULT_MJSOUL, 6)])
def test_cooperative_deal_matches_oracle(mode, rule, n):
    """init_round_coop — one warp deals one round (ChaCha blocks, shuffle, deal, rank sort, wait sets, start_kyoku spread over
    the lanes) — leaves the record init_round leaves: full record after every step of whole games, 4P and sanma."""
    for seed in range(900 + 10 * mode, 900 + 10 * mode + n):
        o, h = OracleBackend(mode, seed, rule), HostsimBackend(mode, seed, rule)
        o.reset()
        h.reset()
        steps = 0
        while True:
            so, sh = o.get_state(), h.get_state()
            d = A.state_fields_equal(so, sh)
            assert not d, f"seed {seed} step {steps}: state differs in {d}"
            if so.is_done:
                break
            o.random_step(0xD1CE, seed)
            h.random_step_coopdeal(0xD1CE, seed)
            steps += 1
        assert o.events() == h.events() and so.kyoku_count > (1 if mode not in (0, 3) else 0)


@pytest.mark.parametrize("mode", [2, 5])
def test_deferred_visits_match_oracle(mode):
    """The rollout kernels park the follow-up of some discards (pending_tail) and the next round's deal (pending_init)
    and run them on later scheduler visits; whenever nothing is parked the record must equal the oracle's at the same
    env-step count, and the event stream must be identical at the end."""
    n_games = 0
    for seed in range(300 * mode, 300 * mode + 10):
        o, h = OracleBackend(mode, seed, A.RULE_DEFAULT_TENHOU), HostsimBackend(mode, seed, A.RULE_DEFAULT_TENHOU)
        o.reset()
        h.reset()
        parked_seen = 0
        for _ in range(20000):
            if h.visit_deferred(0xABCD, seed):
                o.random_step(0xABCD, seed)
            sh = h.get_state()
            if sh.pending_tail[0] != 255 or sh.pending_init[0] != 255:
                parked_seen += 1
                continue
            so = o.get_state()
            d = A.state_fields_equal(so, sh)
            assert not d, f"seed {seed}: state differs in {d}"
            if so.is_done:
                break
        assert o.get_state().is_done and parked_seen > (20 if mode == 5 else 50)
        assert o.events() == h.events()
        n_games += 1
    assert n_games == 10


def test_observation_encode_lockstep(libs):
    """encode() (74x34 f32) and mask() (82) of every acting seat, at every step of seeded games: bit-equal."""
    import numpy as np

    orc, hs = libs
    n_obs = 0
    for seed in (11, 12, 13):
        o, h = OracleBackend(2, seed), HostsimBackend(2, seed)
        o.reset()
        h.reset()
        a = np.zeros(74 * 34, np.float32)
        b = np.zeros(74 * 34, np.float32)
        ma = np.zeros(82, np.uint8)
        mb = np.zeros(82, np.uint8)
        fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        up = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint8))
        step = 0
        while True:
            s = o.get_state()
            if s.is_done:
                break
            for p in range(4):
                if (s.active_mask >> p) & 1:
                    orc.orc_game_encode(o.h, p, fp(a), up(ma))
                    hs.hs_game_encode(h.h, p, fp(b), up(mb))
                    assert a.tobytes() == b.tobytes(), f"seed {seed} step {step} seat {p}: channels {sorted(set(np.nonzero(a != b)[0] // 34))}"
                    assert ma.tobytes() == mb.tobytes(), f"seed {seed} step {step} seat {p}: mask"
                    n_obs += 1
            o.random_step(5, seed)
            h.random_step(5, seed)
            step += 1
    assert n_obs > 3000


def test_observation_encode_extended_lockstep(libs):
    """encode_extended() (215x34 f32) of every acting seat — and of every OTHER seat every 16th step (empty legal list,
    13-tile hands) — at every step of seeded hanchan: bit-equal (the decay rows use the same libm expf on both sides)."""
    import numpy as np

    orc, hs = libs
    n_obs = 0
    for seed in (21, 22):
        o, h = OracleBackend(2, seed), HostsimBackend(2, seed)
        o.reset()
        h.reset()
        a = np.full(215 * 34 + 8, 7.0, np.float32)
        b = np.full(215 * 34 + 8, 7.0, np.float32)
        fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        step = 0
        while True:
            s = o.get_state()
            if s.is_done:
                break
            for p in range(4):
                if (s.active_mask >> p) & 1 or step % 16 == 0:
                    orc.orc_game_encode_ext(o.h, p, fp(a))
                    hs.hs_game_encode_ext(h.h, p, fp(b))
                    assert a.tobytes() == b.tobytes(), f"seed {seed} step {step} seat {p}: channels {sorted(set(np.nonzero(a != b)[0] // 34))}"
                    n_obs += 1
            o.random_step(5, seed)
            h.random_step(5, seed)
            step += 1
    assert n_obs > 2000


def test_kawa_overview_lockstep(libs):
    """encode_kawa_overview() (4x7x34 f32, observer independent) at every step of a seeded hanchan: bit-equal."""
    import numpy as np

    orc, hs = libs
    o, h = OracleBackend(2, 31), HostsimBackend(2, 31)
    o.reset()
    h.reset()
    a = np.full(4 * 7 * 34 + 8, 7.0, np.float32)
    b = np.full(4 * 7 * 34 + 8, 7.0, np.float32)
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    n = 0
    nonzero = 0
    while not o.get_state().is_done:
        orc.orc_game_encode_kawa(o.h, fp(a))
        hs.hs_game_encode_kawa(h.h, fp(b))
        assert a.tobytes() == b.tobytes(), f"step {n}"
        nonzero = max(nonzero, int(a[:4 * 7 * 34].sum()))
        o.random_step(5, 31)
        h.random_step(5, 31)
        n += 1
    assert n > 500 and nonzero > 40


def test_shanten_3p_kernel_source_vs_oracle(libs):
    """shanten_counts_3p (hand.cuh) against the oracle on the 3P golden hands and on random 4-14 tile hands."""
    import os
    import random

    orc, hs = libs
    rng = random.Random(7)
    hands = []
    for line in open(os.path.join(os.path.dirname(__file__), "golden", "shanten3p_golden.txt")):
        if not line.startswith("#"):
            hands.append([int(c) for c in line.split()[0]])
    for _ in range(3000):
        cnt = [0] * 34
        for t in rng.sample(range(136), rng.randrange(4, 15)):
            cnt[t // 4] += 1
        hands.append(cnt)
    for cnt in hands:
        a = (C.c_uint8 * 34)(*cnt)
        assert orc.orc_shanten_counts_3p(a, sum(cnt) // 3) == hs.hs_shanten_counts_3p(a, sum(cnt) // 3), cnt
        assert orc.orc_shanten_counts(a, sum(cnt) // 3) == hs.hs_shanten_counts(a, sum(cnt) // 3), cnt


def test_observation_encode_extended_lockstep_3p(libs):
    """Sanma: Observation3P.encode_extended() (215x27 f32) of every acting seat — and of every seat every 16th step — over a
    seeded sanma hanchan: the scalar definitions of obs_ext3.cuh against the oracle, bit-equal."""
    import numpy as np

    orc, hs = libs
    n_obs = 0
    for seed in (51, 52):
        o, h = OracleBackend(5, seed), HostsimBackend(5, seed)
        o.reset()
        h.reset()
        a = np.full(215 * 27 + 8, 7.0, np.float32)
        b = np.full(215 * 27 + 8, 7.0, np.float32)
        fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        step = 0
        while True:
            s = o.get_state()
            if s.is_done:
                break
            for p in range(3):
                if (s.active_mask >> p) & 1 or step % 16 == 0:
                    orc.orc_game_encode_ext(o.h, p, fp(a))
                    hs.hs_game_encode_ext(h.h, p, fp(b))
                    assert a.tobytes() == b.tobytes(), f"seed {seed} step {step} seat {p}: channels {sorted(set(np.nonzero(a != b)[0] // 27))}"
                    n_obs += 1
            o.random_step(5, seed)
            h.random_step(5, seed)
            step += 1
    assert n_obs > 1000


def test_observation_encode_lockstep_3p(libs):
    """Sanma: Observation3P.encode() (74x27 f32) and mask() (60 ids) of every acting seat at every step: bit-equal."""
    import numpy as np

    orc, hs = libs
    n_obs, kita_ids = 0, 0
    for mode, seed in ((5, 21), (5, 22), (3, 23), (4, 24)):
        o, h = OracleBackend(mode, seed), HostsimBackend(mode, seed)
        o.reset()
        h.reset()
        a = np.full(74 * 27 + 8, 7.0, np.float32)    # guard words: the encoders must write exactly 74*27 floats / 60 bytes
        b = np.full(74 * 27 + 8, 7.0, np.float32)
        ma = np.full(64, 9, np.uint8)
        mb = np.full(64, 9, np.uint8)
        fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
        up = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint8))
        step = 0
        while True:
            s = o.get_state()
            if s.is_done:
                break
            for p in range(3):
                if (s.active_mask >> p) & 1:
                    orc.orc_game_encode(o.h, p, fp(a), up(ma))
                    hs.hs_game_encode(h.h, p, fp(b), up(mb))
                    assert (a[74 * 27:] == 7.0).all() and (b[74 * 27:] == 7.0).all() and (ma[60:] == 9).all() and (mb[60:] == 9).all()
                    assert a.tobytes() == b.tobytes(), f"seed {seed} step {step} seat {p}: channels {sorted(set(np.nonzero(a != b)[0] // 27))}"
                    assert ma.tobytes() == mb.tobytes(), f"seed {seed} step {step} seat {p}: mask"
                    kita_ids += int(ma[59])
                    n_obs += 1
            o.random_step(5, seed)
            h.random_step(5, seed)
            step += 1
    assert n_obs > 1500 and kita_ids > 20


def test_sequence_features_lockstep(libs):
    """encode_seq_{sparse,numeric,progression,candidates} of every acting seat at every step, over the seat's event delta
    since its previous observation (state/mod.rs:211-218): kernel source vs oracle, byte-equal."""
    import numpy as np

    orc, hs = libs
    u16 = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint16))
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    n_obs = n_prog = n_calls = 0
    for seed in (21, 22, 23, 24):
        o, h = OracleBackend(2, seed), HostsimBackend(2, seed)
        o.reset()
        h.reset()
        cursor = [0, 0, 0, 0]
        outs = []
        for _ in range(2):
            outs.append((np.zeros(25, np.uint16), np.zeros(12, np.float32), np.zeros(512 * 5, np.uint16), np.zeros(64 * 4, np.uint16),
                         np.zeros(3, np.uint16)))
        while True:
            s = o.get_state()
            if s.is_done:
                break
            end = s.ev_words
            for p in range(4):
                if (s.active_mask >> p) & 1:
                    for lib_, fn, hnd, out in ((orc, "orc_game_encode_seq", o.h, outs[0]), (hs, "hs_game_encode_seq", h.h, outs[1])):
                        getattr(lib_, fn)(hnd, p, cursor[p], end, 1, u16(out[0]), fp(out[1]), u16(out[2]), 512, u16(out[3]), u16(out[4]))
                    for k, name in enumerate(("sparse", "numeric", "progression", "candidates", "lens")):
                        assert outs[0][k].tobytes() == outs[1][k].tobytes(), f"seed {seed} seat {p} step {s.step_count}: {name}"
                    cursor[p] = end          # the delta advances on every observation
                    n_obs += 1
                    n_prog += int(outs[0][4][1])
                    n_calls += int((outs[0][2].reshape(512, 5)[: outs[0][4][1], 1] >= 38).sum())
                    assert 5 <= outs[0][4][0] <= 25 and outs[0][4][2] >= 1
            o.random_step(9, seed)
            h.random_step(9, seed)
    assert n_obs > 4000 and n_prog > 10000 and n_calls > 100
