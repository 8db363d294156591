"""Parity tests proper: the CUDA library, called through its C ABI, against the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle
from riichienv_b200 import _abi as A
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    return oracle.load()


@pytest.fixture(scope="module")
def ctx():
    from riichienv_b200._lib import Context

    return Context.get(0)


def gpu_eval(ctx, arr, n):
    from riichienv_b200._lib import check, lib

    out = (A.HandResult * n)()
    check(lib().rv_hand_eval_batch(ctx.handle, arr, out, n))
    return out


def test_hand_eval_golden(ctx, orc):
    cases = H.load_agari_cases()
    arr = H.query_array([c[0] for c in cases])
    out = gpu_eval(ctx, arr, len(cases))
    ref = (A.HandResult * len(cases))()
    orc.orc_hand_eval(arr, ref, len(cases))
    for i, (_, exp, yaku) in enumerate(cases):
        assert (out[i].is_win, out[i].han, out[i].fu) == exp, f"case {i}"
        assert H.yaku_ids(out[i].yaku_mask) == yaku, f"case {i}"
        assert bytes(out[i]) == bytes(ref[i]), f"case {i}"


def test_hand_eval_golden_3p(ctx, orc):
    import os

    cases = H.load_agari_cases("agari_3p.txt")
    lines = [l for l in open(os.path.join(H.GOLDEN, "agari_3p.txt")) if not l.startswith("#")]
    for (q, _, _), line in zip(cases, lines):
        q.sanma, q.kita_count = 1, [int(x) for x in line.split("|")[5].split()][4]
    arr = H.query_array([c[0] for c in cases])
    out = gpu_eval(ctx, arr, len(cases))
    ref = (A.HandResult * len(cases))()
    orc.orc_hand_eval(arr, ref, len(cases))
    for i, (_, exp, yaku) in enumerate(cases):
        assert (out[i].is_win, out[i].han, out[i].fu) == exp and H.yaku_ids(out[i].yaku_mask) == yaku, f"case {i}"
        assert bytes(out[i]) == bytes(ref[i]), f"case {i}"


def test_hand_eval_random_vs_oracle(ctx, orc):
    n = 200_000
    qs = H.random_hand_queries(n, seed=11)
    arr = H.query_array(qs)
    out = gpu_eval(ctx, arr, n)
    ref = (A.HandResult * n)()
    orc.orc_hand_eval_mt(arr, ref, n, 16)
    a = np.frombuffer(out, dtype=np.uint8).reshape(n, C.sizeof(A.HandResult))
    b = np.frombuffer(ref, dtype=np.uint8).reshape(n, C.sizeof(A.HandResult))
    bad = np.nonzero((a != b).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} hands differ, first {bad[:5]}"
    assert int(a[:, 28].sum()) > 10_000  # is_win column: positive stratum exercised


def test_hand_eval_config2_full_stream(ctx, orc):
    """BASELINE.json configs[1] at its stated size: the 10^7 DISTINCT seeded hands of include/rv_synth.h, generated on the device
    (rv_hand_queries_seeded), evaluated by the two-kernel path (rv_hand_eval_batch_device) and compared byte for byte with the
    oracle on the same stream (orc_hand_queries_seeded + orc_hand_eval_mt), in chunks of 10^6."""
    import os

    import torch

    from riichienv_b200._lib import check, lib

    total, chunk = 10_000_000, 1_000_000
    qs, rs = C.sizeof(A.HandQuery), C.sizeof(A.HandResult)
    d_q = torch.empty((chunk, qs), dtype=torch.uint8, device="cuda")
    d_r = torch.empty((chunk, rs), dtype=torch.uint8, device="cuda")
    h_q = (A.HandQuery * chunk)()
    h_r = (A.HandResult * chunk)()
    wins = shapes = 0
    for first in range(0, total, chunk):
        check(lib().rv_hand_queries_seeded(ctx.handle, C.c_void_p(d_q.data_ptr()), first, chunk))
        check(lib().rv_hand_eval_batch_device(ctx.handle, C.c_void_p(d_q.data_ptr()), C.c_void_p(d_r.data_ptr()), chunk))
        ctx.sync()
        orc.orc_hand_queries_seeded(h_q, first, chunk)
        assert np.array_equal(d_q.cpu().numpy(), np.frombuffer(h_q, np.uint8).reshape(chunk, qs)), "query streams differ"
        orc.orc_hand_eval_mt(h_q, h_r, chunk, os.cpu_count() or 1)
        a, b = d_r.cpu().numpy(), np.frombuffer(h_r, np.uint8).reshape(chunk, rs)
        bad = np.nonzero((a != b).any(axis=1))[0]
        assert bad.size == 0, f"chunk at {first}: {bad.size} hands differ, first {first + int(bad[0])}"
        wins += int(a[:, 28].sum())
        shapes += int(a[:, 30].sum())
    assert shapes > 0.09 * total and wins > 0.07 * total    # the positive stratum (every tenth hand) reaches yaku / fu / score


def test_shanten_golden_via_hand_eval(ctx):
    cases = H.load_counts_file("shanten_golden.txt")
    qs = []
    keep = []
    for cnt, sh in cases:
        tiles = [t * 4 + k for t in range(34) for k in range(cnt[t])]
        if len(tiles) % 3 != 2 or len(tiles) > 14:
            continue
        qs.append(H.make_query(tiles, [], tiles[-1], [], [], 0, 0, 0, 0))
        keep.append(sh)
    out = gpu_eval(ctx, H.query_array(qs), len(qs))
    assert len(qs) > 1500
    for i, sh in enumerate(keep):
        assert out[i].shanten == sh, f"hand {i}"


def test_shanten_3p_golden_via_hand_eval(ctx):
    """calculate_shanten_3p through rv_hand_eval_batch (sanma queries): the reference tables' answers of
    tests/golden/shanten3p_golden.txt, and the shim function on a known answer of tests/test_shanten.py."""
    cases = H.load_counts_file("shanten3p_golden.txt")
    qs = []
    keep = []
    for cnt, sh in cases:
        tiles = [t * 4 + k for t in range(34) for k in range(cnt[t])]
        if len(tiles) % 3 != 2 or len(tiles) > 14:
            continue
        q = H.make_query(tiles, [], tiles[-1], [], [], 0, 0, 0, 0)
        q.sanma = 1
        qs.append(q)
        keep.append(sh)
    out = gpu_eval(ctx, H.query_array(qs), len(qs))
    assert len(qs) > 700
    for i, sh in enumerate(keep):
        assert out[i].shanten == sh, f"hand {i}"
    from riichienv_b200 import calculate_shanten, calculate_shanten_3p

    hand = [0, 1, 2, 3, 108, 109, 110, 111, 112, 113, 114, 116, 117]      # 1111m111122233z (tests/test_shanten.py:4-13)
    assert calculate_shanten(hand) == 1 and calculate_shanten_3p(hand) == 2


def run_both(orc, n, mode, rule, seed_base, agent_seed, max_steps=100000):
    from riichienv_b200.vec_env import VecRiichiEnv

    v = VecRiichiEnv(n, mode, rule, seed_base=seed_base)
    v.reset()
    total = v.step_random(agent_seed, max_steps)
    done, scores, ranks = v.results()
    sc, kc, ec, eh = v.counters()
    v.close()
    o_scores = np.zeros((n, 4), np.int32)
    o_ranks = np.zeros((n, 4), np.uint8)
    o_done = np.zeros(n, np.uint8)
    o_steps = np.zeros(n, np.uint32)
    o_ky = np.zeros(n, np.uint32)
    o_ec = np.zeros(n, np.uint32)
    o_h = np.zeros(n, np.uint64)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    o_total = orc.orc_run_random(mode, rule, seed_base, n, agent_seed, max_steps, 32, p(o_scores, C.c_int32), p(o_ranks, C.c_uint8),
                                 p(o_done, C.c_uint8), p(o_steps, C.c_uint32), p(o_ky, C.c_uint32), p(o_ec, C.c_uint32),
                                 p(o_h, C.c_uint64), None)
    return (total, done, scores, ranks, sc, kc, ec, eh), (o_total, o_done, o_scores, o_ranks, o_steps, o_ky, o_ec, o_h)


@pytest.mark.parametrize("mode,rule,n", [(2, A.RULE_DEFAULT_TENHOU, 4096), (2, A.RULE_DEFAULT_MJSOUL, 1024),
                                          (1, A.RULE_DEFAULT_TENHOU, 1024), (0, A.RULE_DEFAULT_TENHOU, 4096),
                                          (5, A.RULE_DEFAULT_TENHOU, 4096), (5, A.RULE_DEFAULT_MJSOUL, 1024),
                                          (4, A.RULE_DEFAULT_TENHOU, 1024), (3, A.RULE_DEFAULT_TENHOU, 4096)])
def test_random_games_vs_oracle(orc, mode, rule, n):
    g, o = run_both(orc, n, mode, rule, seed_base=1000 * mode, agent_seed=0xABCDEF)
    assert g[0] == o[0], "total env steps"
    names = ["done", "scores", "ranks", "step_count", "kyoku_count", "ev_count", "ev_hash"]
    for name, a, b in zip(names, g[1:], o[1:]):
        assert np.array_equal(a, b), f"{name} differs in {int((a != b).sum())} entries"
    assert g[1].all()   # incl. the rare stalled 3P games, which the rollout retires (overflow bit 1)


@pytest.mark.parametrize("n", [1, 31, 33, 383, 385, 12 * 32 * 148 + 1])
def test_rollout_scheduler_odd_sizes(orc, n):
    """vector sizes around the scheduler's granules (a warp = 32 games, a block = 384, the crew = 148 blocks), incl. step budgets
    that end the call in the middle of a game and a resume"""
    from riichienv_b200.vec_env import VecRiichiEnv

    g, o = run_both(orc, n, 2, A.RULE_DEFAULT_TENHOU, seed_base=31_000 + n, agent_seed=0x0DD)
    assert g[0] == o[0]
    for a, b in zip(g[1:], o[1:]):
        assert np.array_equal(a, b)
    if n <= 385:
        v = VecRiichiEnv(n, 2, A.RULE_DEFAULT_TENHOU, seed_base=31_000 + n)
        v.reset()
        total = 0
        for budget in (40, 333, 100000):                 # three calls: the rollout resumes where the budget stopped it
            total += v.step_random(0x0DD, budget)
        assert total == g[0] and np.array_equal(v.results()[1], g[2]) and np.array_equal(v.counters()[3], g[7])
        v.close()


def test_sanma_config_size_65536_games(orc):
    """BASELINE configs[3] at its stated size: 65,536 sanma hanchan (3p-red-half), done / scores / ranks / counters / event hash."""
    n = 65536
    g, o = run_both(orc, n, 5, A.RULE_DEFAULT_TENHOU, seed_base=6_000_000, agent_seed=0x3A3A)
    assert g[0] == o[0], "total env steps"
    for name, a, b in zip(["done", "scores", "ranks", "step_count", "kyoku_count", "ev_count", "ev_hash"], g[1:], o[1:]):
        assert np.array_equal(a, b), f"{name} differs in {int((a != b).sum())} entries"
    assert g[1].all()


def test_parity_gate_100k_games(orc):
    """The parity gate that accompanies the headline number (SURVEY.md section 8 d): >= 10^5 hanchan, final done / scores /
    ranks / step count / round count / event count / 64-bit event-stream hash equal between the CUDA rollout and the oracle."""
    n = 102400
    g, o = run_both(orc, n, 2, A.RULE_DEFAULT_TENHOU, seed_base=5_000_000, agent_seed=0x5EED)
    assert g[0] == o[0], "total env steps"
    for name, a, b in zip(["done", "scores", "ranks", "step_count", "kyoku_count", "ev_count", "ev_hash"], g[1:], o[1:]):
        assert np.array_equal(a, b), f"{name} differs in {int((a != b).sum())} entries"
    assert g[1].all() and g[0] > 1000 * n


@pytest.mark.gpu
@pytest.mark.parametrize("mode,policy", [(0, 0), (0, 1), (3, 0), (3, 1)])
def test_explicit_wall_gate_100k_rounds(orc, mode, policy):
    """The parity gate WITHOUT the seeded shuffle (the one function whose parity with the reference cannot be pinned here):
    102,400 single-round games dealt from caller-supplied walls (reset(wall=) -> load_wall, state/wall.rs:69-80; numpy
    permutations of the 136 / 108 tile ids), played to the end by the random (policy 0) and the greedy-win agent (policy 1)."""
    from riichienv_b200.vec_env import VecRiichiEnv

    n, seed_base, agent_seed = 102400, 77_000_000, 0xE7A11
    wl = 108 if mode >= 3 else 136
    tiles = np.array([t for t in range(136) if mode < 3 or not (4 <= t < 32)], np.uint8)      # sanma: no 2m-8m
    assert tiles.size == wl
    rng = np.random.default_rng(20260117 + mode)
    walls = np.ascontiguousarray(rng.permuted(np.broadcast_to(tiles, (n, wl)), axis=1))
    v = VecRiichiEnv(n, mode, A.RULE_DEFAULT_TENHOU, seed_base=seed_base)
    v.reset(walls=walls)
    # policy 0 goes through the rollout scheduler (the crew kernel), policy 1 through the agent kernel
    total = v.step_random(agent_seed, 100000) if policy == 0 else v.step_agent(policy, agent_seed, 100000)
    done, scores, ranks = v.results()
    sc, kc, ec, eh = v.counters()
    v.close()
    o_scores, o_ranks, o_done = np.zeros((n, 4), np.int32), np.zeros((n, 4), np.uint8), np.zeros(n, np.uint8)
    o_steps, o_ky, o_ec, o_h = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint64)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    import os

    o_total = orc.orc_run_agent_walls(policy, mode, A.RULE_DEFAULT_TENHOU, seed_base, n, agent_seed, 100000, os.cpu_count() or 1,
                                      p(walls, C.c_uint8), p(o_scores, C.c_int32), p(o_ranks, C.c_uint8), p(o_done, C.c_uint8),
                                      p(o_steps, C.c_uint32), p(o_ky, C.c_uint32), p(o_ec, C.c_uint32), p(o_h, C.c_uint64))
    assert total == o_total and done.all() and o_done.all()
    for name, a, b in zip(["scores", "ranks", "step_count", "kyoku_count", "ev_count", "ev_hash"], [scores, ranks, sc, kc, ec, eh],
                          [o_scores, o_ranks, o_steps, o_ky, o_ec, o_h]):
        assert np.array_equal(a, b), f"{name} differs in {int((a != b).sum())} entries"
    assert (kc == 1).all()                                     # one round each: the seeded shuffle is never reached
    if policy == 1:
        assert (scores[:, : 3 if mode >= 3 else 4] != (35000 if mode >= 3 else 25000)).any(axis=1).mean() > 0.5   # most rounds are won


def run_both_agent(orc, policy, n, mode, rule, seed_base, agent_seed, hist=None, max_steps=200000):
    from riichienv_b200.vec_env import VecRiichiEnv

    v = VecRiichiEnv(n, mode, rule, seed_base=seed_base)
    v.reset()
    total = v.step_agent(policy, agent_seed, max_steps)
    done, scores, ranks = v.results()
    sc, kc, ec, eh = v.counters()
    v.close()
    o_scores, o_ranks, o_done = np.zeros((n, 4), np.int32), np.zeros((n, 4), np.uint8), np.zeros(n, np.uint8)
    o_steps, o_ky, o_ec, o_h = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint64)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    import os

    o_total = orc.orc_run_agent(policy, mode, rule, seed_base, n, agent_seed, max_steps, os.cpu_count() or 1, p(o_scores, C.c_int32),
                                p(o_ranks, C.c_uint8), p(o_done, C.c_uint8), p(o_steps, C.c_uint32), p(o_ky, C.c_uint32),
                                p(o_ec, C.c_uint32), p(o_h, C.c_uint64), None if hist is None else p(hist, C.c_uint64))
    return (total, done, scores, ranks, sc, kc, ec, eh), (o_total, o_done, o_scores, o_ranks, o_steps, o_ky, o_ec, o_h)


@pytest.mark.parametrize("mode,rule,n", [(2, A.RULE_DEFAULT_MJSOUL, 4096), (0, A.RULE_DEFAULT_TENHOU, 8192), (1, A.RULE_DEFAULT_MJSOUL, 2048),
                                          (5, A.RULE_DEFAULT_TENHOU, 4096), (5, A.RULE_DEFAULT_MJSOUL, 4096), (3, A.RULE_DEFAULT_MJSOUL, 8192)])
def test_greedy_agent_games_vs_oracle(orc, mode, rule, n):
    """rv_vec_step_agent(RV_AGENT_GREEDY): every mode and both rule presets, final records and event hashes vs the oracle."""
    g, o = run_both_agent(orc, 1, n, mode, rule, seed_base=7000 * (mode + 1), agent_seed=0xFACE)
    assert g[0] == o[0], "total env steps"
    for name, a, b in zip(["done", "scores", "ranks", "step_count", "kyoku_count", "ev_count", "ev_hash"], g[1:], o[1:]):
        assert np.array_equal(a, b), f"{name} differs in {int((a != b).sum())} entries"
    assert g[1].all()


def test_settlement_gate_100k_greedy_games(orc):
    """A5 (tsumo / ron settlement, state/mod.rs:685-893, 919-1142) at the size of the headline parity gate: 102,400 hanchan
    played by the greedy-win agent — about 60 % of ~1.1 M rounds end in a win — with done / scores / ranks / step, round and
    event counts and the 64-bit hash of the event stream (every hora event carries han, fu, the yaku set, deltas and ura
    markers) equal between the CUDA path and the oracle.  The histogram (from the oracle's logs, which the hashes prove equal
    to the GPU's) asserts that the quirk-laden branches were actually taken."""
    n = 102400
    hist = np.zeros(128, np.uint64)
    g, o = run_both_agent(orc, 1, n, 2, A.RULE_DEFAULT_TENHOU, seed_base=9_000_000, agent_seed=0xA5A5, hist=hist)
    assert g[0] == o[0], "total env steps"
    for name, a, b in zip(["done", "scores", "ranks", "step_count", "kyoku_count", "ev_count", "ev_hash"], g[1:], o[1:]):
        assert np.array_equal(a, b), f"{name} differs in {int((a != b).sum())} entries"
    assert g[1].all()
    H = {k: int(hist[i]) for k, i in dict(hora=64, tsumo=65, ron=66, multi_ron=67, pao=68, rounds=69, ryukyoku=70,
                                          yakuman=80, kazoe=81).items()}
    yaku = {y: int(hist[y]) for y in range(64) if hist[y]}
    print("greedy gate:", H, "yaku:", yaku, "ryukyoku by reason:", [int(x) for x in hist[71:80]])
    assert H["hora"] > 0.3 * H["rounds"], "more than 30 % of the rounds must end in a win"
    assert H["tsumo"] > 10000 and H["ron"] > 10000 and H["multi_ron"] > 100
    # chankan 3, rinshan 4, haitei 5, houtei 6, ippatsu 30, ura 33, double riichi 18; yakuman: at least daisangen / suuankou
    # (tanki) / kokushi (13-wait) families; suucha riichi (reason 5) and sanchaho (reason 6) draws
    for y in (1, 2, 3, 4, 5, 6, 18, 30, 31, 32, 33):
        assert yaku.get(y, 0) > 0, f"yaku {y} never occurred"
    assert yaku.get(37, 0) > 0 and yaku.get(38, 0) + yaku.get(48, 0) > 0 and yaku.get(42, 0) + yaku.get(49, 0) > 0
    assert H["yakuman"] >= 100 and hist[71 + 5] > 0 and hist[71 + 6] > 0


@pytest.mark.parametrize("devices", [(0,), (0, 0, 0), "all"])
def test_multi_device_results_do_not_depend_on_device_count(devices):
    """rv_multi_*: the same 65,536 seeded hanchan as ONE vector and sharded — over three shards on one GPU (exercises the
    partitioning and the per-device host threads on a single-GPU box) and over every GPU of the box.  Scores, ranks and event
    hashes are identical game by game; the statistics reduction adds up."""
    import torch

    from riichienv_b200.vec_env import MultiVecRiichiEnv, VecRiichiEnv

    if devices == "all":
        if torch.cuda.device_count() < 2:
            pytest.skip("needs at least two GPUs")
        devices = tuple(range(torch.cuda.device_count()))
    n, seed_base = 65536, 123_000
    v = VecRiichiEnv(n, 2, A.RULE_DEFAULT_TENHOU, seed_base=seed_base)
    v.reset()
    total = v.step_random(0xABBA, 100000)
    done, scores, ranks = v.results()
    sc, kc, ec, eh = v.counters()
    v.close()
    m = MultiVecRiichiEnv(n, 2, A.RULE_DEFAULT_TENHOU, seed_base=seed_base, devices=devices)
    m.reset()
    m_total = m.step_random(0xABBA, 100000)
    m_done, m_scores, m_ranks = m.results()
    m_sc, m_kc, m_ec, m_eh = m.counters()
    st = m.stats()
    first = [m.shard(k)[1] for k in range(len(devices))]
    m.close()
    assert first == [n * k // len(devices) for k in range(len(devices))]
    assert m_total == total and done.all() and m_done.all()
    for name, a, b in (("scores", scores, m_scores), ("ranks", ranks, m_ranks), ("steps", sc, m_sc), ("rounds", kc, m_kc),
                       ("events", ec, m_ec), ("hash", eh, m_eh)):
        assert np.array_equal(a, b), f"{name} differ in {int((a != b).sum())} entries"
    assert st["games"] == n and st["games_done"] == n and st["env_steps"] == total and st["rounds"] == int(kc.sum())
    assert st["score_sum"] == [int(x) for x in scores.sum(0)]
    assert sum(st["rank_hist"][0]) == n and st["rank_hist"][2][0] == int((ranks[:, 2] == 1).sum())


def test_scheduler_watchdog_reports_instead_of_hanging():
    """The persistent rollout ends when its live-game counter reaches zero.  With one more live game than the queues hold
    (fault injection) the crew can never finish: the watchdog must flag the launch — rv_vec_steps_total returns RV_ERR_CUDA —
    within seconds instead of hanging the GPU, and the next rollout on the same vector must be clean."""
    import os
    import time

    from riichienv_b200._lib import RvError
    from riichienv_b200.vec_env import VecRiichiEnv

    v = VecRiichiEnv(2048, 2, A.RULE_DEFAULT_TENHOU, seed_base=4242)
    v.reset()
    os.environ["RV_FAULT_INJECT"] = "lost_game"
    try:
        t0 = time.time()
        with pytest.raises(RvError, match="watchdog"):
            v.step_random(3, 100000)
        assert time.time() - t0 < 60
    finally:
        del os.environ["RV_FAULT_INJECT"]
    done, _, _ = v.results()
    assert done.all()                       # the games themselves were played out; only the termination was sabotaged
    v.reset()
    total = v.step_random(3, 100000)
    done, scores, _ = v.results()
    assert done.all() and total > 2048 * 500
    v.close()


def test_partial_rollout_and_resume(orc):
    """max_steps < game length: state must carry over between launches exactly."""
    from riichienv_b200.vec_env import VecRiichiEnv

    n = 512
    v = VecRiichiEnv(n, 2, A.RULE_DEFAULT_TENHOU, seed_base=77)
    v.reset()
    total = 0
    for _ in range(40):
        total += v.step_random(5, 100)
    done, scores, ranks = v.results()
    sc, kc, ec, eh = v.counters()
    o_scores = np.zeros((n, 4), np.int32)
    o_h = np.zeros(n, np.uint64)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    o_total = orc.orc_run_random(2, A.RULE_DEFAULT_TENHOU, 77, n, 5, 4000, 32, p(o_scores, C.c_int32), None, None, None, None, None,
                                 p(o_h, C.c_uint64), None)
    assert total == o_total and np.array_equal(scores, o_scores) and np.array_equal(eh, o_h)


def test_lockstep_snapshots_legal_and_events(orc):
    """Per-step: full state record, legal-action lists; at the end: event log words and MJAI JSON."""
    from tests.backends import GpuBackend, OracleBackend

    for mode, seed in ((2, 3), (2, 4), (5, 6), (3, 7)):
        g, o = GpuBackend(mode, seed), OracleBackend(mode, seed)
        g.reset()
        o.reset()
        steps = 0
        while True:
            sg, so = g.get_state(), o.get_state()
            d = A.state_fields_equal(so, sg)
            assert not d, f"seed {seed} step {steps}: {d}"
            if so.is_done:
                break
            if steps % 7 == 0:
                for p in range(4):
                    assert g.legal_tuples(p) == o.legal_tuples(p)
            g.random_step(99, seed)
            o.random_step(99, seed)
            steps += 1
        assert g.events() == o.events()
        for viewer in [-1] + list(range(3 if mode >= 3 else 4)):     # the all-seeing log and every seat's masked view:
            assert g.events_json(viewer) == o.events_json(viewer)    # product renderer vs the oracle's own text log


def test_lockstep_legal_lists_1024_games(orc):
    """The parity gate of SURVEY 8(d) on 1,024 hanchan driven in lock-step: the legal-action list of every seat at EVERY
    decision point (rv_vec_legal_actions vs the oracle, entries and order), then the complete binary event log of every game
    (what the MJAI JSON is rendered from), final scores and ranks."""
    from riichienv_b200._lib import events_to_json
    from riichienv_b200.vec_env import VecRiichiEnv

    n, seed_base, agent = 1024, 31000, 17
    v = VecRiichiEnv(n, 2, A.RULE_DEFAULT_TENHOU, seed_base=seed_base, log_cap_words=1 << 14)
    v.reset()
    hs = (C.c_void_p * n)(*[orc.orc_game_new(2, seed_base + g, 0, A.RULE_DEFAULT_TENHOU, 1) for g in range(n)])
    for h in hs:
        orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    o_acts = (A.Action * (n * 4 * A.MAX_LEGAL))()
    o_cnt = np.zeros((n, 4), np.uint8)
    k_idx = np.arange(A.MAX_LEGAL)[None, None, :]
    steps = lists = 0
    while True:
        g_acts, g_cnt = v.legal_actions()
        orc.orc_games_legal_batch(hs, n, o_acts, o_cnt.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert np.array_equal(g_cnt, o_cnt), f"step {steps}: legal-list lengths differ in games {np.nonzero((g_cnt != o_cnt).any(1))[0][:8]}"
        if not g_cnt.any():
            break
        ga = np.frombuffer(g_acts, np.uint8).reshape(n, 4, A.MAX_LEGAL, C.sizeof(A.Action))
        oa = np.frombuffer(o_acts, np.uint8).reshape(n, 4, A.MAX_LEGAL, C.sizeof(A.Action))
        live = (k_idx < g_cnt[:, :, None])[..., None]
        bad = ((ga != oa) & live).any(axis=(1, 2, 3))
        assert not bad.any(), f"step {steps}: legal lists differ in games {np.nonzero(bad)[0][:8]}"
        lists += int((g_cnt > 0).sum())
        v.step_random(agent, 1)
        orc.orc_games_random_step_batch(hs, n, agent, seed_base)
        steps += 1
        assert steps < 6000
    done, scores, ranks = v.results()
    assert done.all() and lists > 900_000
    buf = (C.c_uint32 * (1 << 14))()
    st = A.GameState()
    for g in range(n):
        words = v.events(g)
        nw = orc.orc_game_events(hs[g], buf, 1 << 14)
        assert nw == len(words) and list(buf[:nw]) == words, f"game {g}: event logs differ"
        orc.orc_game_snapshot(hs[g], C.byref(st))
        assert [st.score[p] for p in range(4)] == list(scores[g])
        # full MJAI text of every game: the product's renderer against the text the oracle wrote at event time
        # (oracle/json.hpp), the all-seeing log of every game and the four masked views of every 16th
        for viewer in ([-1, 0, 1, 2, 3] if g % 16 == 0 else [-1]):
            ln = orc.orc_game_mjai_log(hs[g], viewer, None, 0)
            tb = C.create_string_buffer(ln + 1)
            orc.orc_game_mjai_log(hs[g], viewer, tb, ln + 1)
            js = events_to_json(words, viewer)
            assert js == tb.value.decode().split("\n"), f"game {g} viewer {viewer}: MJAI text differs"
            assert js[0] == '{"type":"start_game"}' and js[-1] == '{"type":"end_game"}'
    for h in hs:
        orc.orc_game_free(h)


def test_external_actions_step(orc):
    """rv_vec_step with host-chosen actions (legal list -> keyed pick on the host) matches the on-device agent."""
    from riichienv_b200.vec_env import VecRiichiEnv

    n = 64
    a = VecRiichiEnv(n, 0, A.RULE_DEFAULT_TENHOU, seed_base=500)
    b = VecRiichiEnv(n, 0, A.RULE_DEFAULT_TENHOU, seed_base=500)
    a.reset()
    b.reset()
    mix = lambda z: ((((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9 % 2**64) ^ ((((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) % 2**64) >> 27)) * 0x94D049BB133111EB) % 2**64
    def mix64(z):
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) % 2**64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) % 2**64
        return z ^ (z >> 31)
    for it in range(400):
        done, _, _ = a.results()
        if done.all():
            break
        acts, counts = a.legal_actions()
        sc, _, _, _ = a.counters()
        chosen = (A.Action * (n * 4))()
        for g in range(n):
            for p in range(4):
                k = int(counts[g, p])
                if k == 0:
                    chosen[g * 4 + p].type = A.NO_ACTION
                    continue
                key = (0x1234 ^ (((500 + g) * 0x9E3779B97F4A7C15) % 2**64) ^ (int(sc[g]) << 8) ^ p) % 2**64
                chosen[g * 4 + p] = acts[(g * 4 + p) * A.MAX_LEGAL + mix64(key) % k]
        a.step(chosen)
        b.step_random(0x1234, 1)
    ra, rb = a.results(), b.results()
    for x, y in zip(ra, rb):
        assert np.array_equal(x, y)
    assert np.array_equal(a.counters()[3], b.counters()[3])


@pytest.mark.parametrize("mode,impl", [(2, "queue"), (5, "queue"), (2, "fused"), (5, "fused")])
def test_observe_step_fused_vs_oracle(orc, mode, impl, monkeypatch):
    """rv_vec_observe_step_random — both implementations (queue: encode kernel + one-step class-queue rollout + mask rows;
    fused: the single kernel): at every step of 192 games the tensor and mask rows equal the oracle's encode()/mask()
    of the state BEFORE the step, and the games end with the oracle's scores and event hashes."""
    import torch

    monkeypatch.setenv("RV_OBS_STEP", impl)

    from riichienv_b200.vec_env import VecRiichiEnv

    n = 192
    W, IDS = (27, 60) if mode >= 3 else (34, 82)
    v = VecRiichiEnv(n, mode, A.RULE_DEFAULT_TENHOU, seed_base=7700)
    v.reset()
    games = [orc.orc_game_new(mode, 7700 + g, 0, A.RULE_DEFAULT_TENHOU, 0) for g in range(n)]
    for h in games:
        orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    obs = torch.empty((n * 3, 74, W), dtype=torch.float32, device="cuda")
    mask = torch.empty((n * 3, IDS), dtype=torch.uint8, device="cuda")
    idx = torch.empty((n * 3,), dtype=torch.int32, device="cuda")
    a = np.zeros(74 * W, np.float32)
    m = np.zeros(IDS, np.uint8)
    st = A.GameState()
    checked = 0
    for it in range(400):
        check_now = it < 40 or it % 7 == 0
        obs.fill_(-1.0)
        mask.fill_(7)
        rows = v.observe_step_random(33, obs=obs, mask=mask, index=idx, sync=True)
        if check_now:
            h_obs, h_mask, h_idx = obs[:rows].cpu().numpy(), mask[:rows].cpu().numpy(), idx[:rows].cpu().numpy()
            r = 0
            for g in range(n):
                orc.orc_game_snapshot(games[g], C.byref(st))
                if st.is_done:
                    continue
                for p in range(4):
                    if (st.active_mask >> p) & 1:
                        assert h_idx[r] == g * 4 + p
                        orc.orc_game_encode(games[g], p, a.ctypes.data_as(C.POINTER(C.c_float)), m.ctypes.data_as(C.POINTER(C.c_uint8)))
                        assert h_obs[r].tobytes() == a.tobytes(), f"iter {it} game {g} seat {p}: channels {sorted(set(np.nonzero(h_obs[r].ravel() != a)[0] // W))}"
                        assert h_mask[r].tobytes() == m.tobytes(), f"iter {it} game {g} seat {p} mask"
                        r += 1
                        checked += 1
            assert r == rows
            assert (obs[rows:] == -1.0).all() and (mask[rows:] == 7).all()     # nothing written past the last row
        for g in range(n):
            orc.orc_game_random_step(games[g], 33, 7700 + g)
    # run both to the end: same results
    while True:
        rows = v.observe_step_random(33, obs=obs, mask=mask, index=idx, sync=True)
        if rows == 0:
            break
    done, scores, _ = v.results()
    _, _, _, eh = v.counters()
    assert done.all()
    for g in range(n):
        while True:
            orc.orc_game_snapshot(games[g], C.byref(st))
            if st.is_done:
                break
            orc.orc_game_random_step(games[g], 33, 7700 + g)
        assert [st.score[p] for p in range(4)] == list(scores[g]) and st.ev_hash == eh[g], f"game {g}"
        orc.orc_game_free(games[g])
    assert checked > 5000


@pytest.mark.parametrize("mode", [2, 5])
def test_observation_encode_vs_oracle(orc, mode):
    """rv_vec_encode: every acting seat of 256 games at several points of the rollout, bytes equal to the oracle's
    Observation::encode / mask restatement (4P: 74x34 + 82 ids; sanma: 74x27 + 60 ids); row order ascending (game, seat)."""
    import torch

    from riichienv_b200.vec_env import VecRiichiEnv

    n = 256
    W, IDS = (27, 60) if mode >= 3 else (34, 82)
    v = VecRiichiEnv(n, mode, A.RULE_DEFAULT_TENHOU, seed_base=9000)
    v.reset()
    games = [orc.orc_game_new(mode, 9000 + g, 0, A.RULE_DEFAULT_TENHOU, 0) for g in range(n)]
    for h in games:
        orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    obs = torch.empty((n * 3, 74, W), dtype=torch.float32, device="cuda")
    mask = torch.empty((n * 3, IDS), dtype=torch.uint8, device="cuda")
    idx = torch.empty((n * 3,), dtype=torch.int32, device="cuda")
    a = np.zeros(74 * W, np.float32)
    m = np.zeros(IDS, np.uint8)
    checked = 0
    for it in range(60):
        rows = v.encode(obs=obs, mask=mask, index=idx)
        h_obs, h_mask, h_idx = obs[:rows].cpu().numpy(), mask[:rows].cpu().numpy(), idx[:rows].cpu().numpy()
        assert (np.diff(h_idx) > 0).all()
        st = A.GameState()
        expect_rows = 0
        for g in range(n):
            orc.orc_game_snapshot(games[g], C.byref(st))
            if st.is_done:
                continue
            for p in range(4):
                if (st.active_mask >> p) & 1:
                    assert h_idx[expect_rows] == g * 4 + p
                    orc.orc_game_encode(games[g], p, a.ctypes.data_as(C.POINTER(C.c_float)), m.ctypes.data_as(C.POINTER(C.c_uint8)))
                    assert h_obs[expect_rows].tobytes() == a.tobytes(), f"iter {it} game {g} seat {p}"
                    assert h_mask[expect_rows].tobytes() == m.tobytes(), f"iter {it} game {g} seat {p} mask"
                    expect_rows += 1
                    checked += 1
        assert expect_rows == rows
        stride = 1 if it < 30 else 37
        v.step_random(21, stride)
        for g in range(n):
            for _ in range(stride):
                orc.orc_game_random_step(games[g], 21, 9000 + g)
    for h in games:
        orc.orc_game_free(h)
    assert checked > (5000 if mode >= 3 else 10000)


@pytest.mark.parametrize("mode", [2, 5])
def test_observation_encode_100k_rows(orc, mode):
    """SURVEY §8 d: encode() bytes equal on >= 10^5 observations.  2,048 hanchan, every acting seat at 64 decision points
    spread over the rollout (rv_vec_encode vs the oracle's encode / mask restatement, whole buffers compared at once)."""
    import torch

    from riichienv_b200.vec_env import VecRiichiEnv

    n, seed_base, agent = (4096 if mode >= 3 else 2048), 9100, 71      # sanma hanchan are shorter: fewer rows per game
    W, IDS = (27, 60) if mode >= 3 else (34, 82)
    v = VecRiichiEnv(n, mode, A.RULE_DEFAULT_TENHOU, seed_base=seed_base)
    v.reset()
    hs = (C.c_void_p * n)(*[orc.orc_game_new(mode, seed_base + g, 0, A.RULE_DEFAULT_TENHOU, 0) for g in range(n)])
    for h in hs:
        orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    cap = n * 3
    obs = torch.empty((cap, 74, W), dtype=torch.float32, device="cuda")
    mask = torch.empty((cap, IDS), dtype=torch.uint8, device="cuda")
    idx = torch.empty((cap,), dtype=torch.int32, device="cuda")
    o_obs, o_mask, o_idx = np.zeros((cap, 74, W), np.float32), np.zeros((cap, IDS), np.uint8), np.zeros(cap, np.int32)
    p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    total = 0
    for point in range(64):
        rows = v.encode(obs=obs, mask=mask, index=idx)
        o_rows = orc.orc_games_encode_batch(hs, n, p(o_obs, C.c_float), p(o_mask, C.c_uint8), p(o_idx, C.c_int32), cap)
        assert rows == o_rows, f"point {point}: {rows} rows vs {o_rows}"
        if rows == 0:
            break
        assert np.array_equal(idx[:rows].cpu().numpy(), o_idx[:rows])
        g_obs = obs[:rows].cpu().numpy()
        bad = np.nonzero((g_obs.view(np.uint32) != o_obs[:rows].view(np.uint32)).any(axis=(1, 2)))[0]
        assert bad.size == 0, f"point {point}: {bad.size} tensor rows differ, first row {int(bad[0])} (game*4+seat {int(o_idx[bad[0]])})"
        assert np.array_equal(mask[:rows].cpu().numpy(), o_mask[:rows]), f"point {point}: mask rows differ"
        total += rows
        steps = 1 if point < 8 else 19             # the first turns one by one, then strides through the hanchan
        v.step_random(agent, steps)
        for _ in range(steps):
            orc.orc_games_random_step_batch(hs, n, agent, seed_base)
    assert total >= 100_000, total
    for h in hs:
        orc.orc_game_free(h)


@pytest.mark.parametrize("mode", [2, 5])
def test_observation_encode_extended_vs_oracle(orc, mode):
    """rv_vec_encode_ext: Observation::encode_extended (215x34; sanma 215x27) + mask of every acting seat of 192 hanchan at many
    points of the rollout, bytes equal to the oracle's restatement (observation/encode.rs:12-584, observation_3p/encode.rs:22-620);
    nothing written past the last row."""
    import torch

    from riichienv_b200.vec_env import VecRiichiEnv

    n = 192
    W, IDS = (27, 60) if mode >= 3 else (34, 82)
    v = VecRiichiEnv(n, mode, A.RULE_DEFAULT_TENHOU, seed_base=9500)
    v.reset()
    games = [orc.orc_game_new(mode, 9500 + g, 0, A.RULE_DEFAULT_TENHOU, 0) for g in range(n)]
    for h in games:
        orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    obs = torch.empty((n * 3, 215, W), dtype=torch.float32, device="cuda")
    mask = torch.empty((n * 3, IDS), dtype=torch.uint8, device="cuda")
    idx = torch.empty((n * 3,), dtype=torch.int32, device="cuda")
    a = np.zeros(215 * W, np.float32)
    m = np.zeros(IDS, np.uint8)
    checked = 0
    seen_channels = np.zeros(215, bool)
    for it in range(50):
        obs.fill_(-3.0)
        rows = v.encode_extended(obs=obs, mask=mask, index=idx)
        assert (obs[rows:] == -3.0).all()
        h_obs, h_mask, h_idx = obs[:rows].cpu().numpy(), mask[:rows].cpu().numpy(), idx[:rows].cpu().numpy()
        assert (np.diff(h_idx) > 0).all()
        st = A.GameState()
        expect_rows = 0
        for g in range(n):
            orc.orc_game_snapshot(games[g], C.byref(st))
            if st.is_done:
                continue
            for p in range(4):
                if (st.active_mask >> p) & 1:
                    assert h_idx[expect_rows] == g * 4 + p
                    orc.orc_game_encode_ext(games[g], p, a.ctypes.data_as(C.POINTER(C.c_float)))
                    orc.orc_game_encode(games[g], p, None, m.ctypes.data_as(C.POINTER(C.c_uint8)))
                    if h_obs[expect_rows].tobytes() != a.tobytes():
                        bad = sorted(set(np.nonzero(h_obs[expect_rows].ravel() != a)[0] // W))
                        raise AssertionError(f"iter {it} game {g} seat {p}: channels {bad}")
                    assert h_mask[expect_rows].tobytes() == m.tobytes(), f"iter {it} game {g} seat {p} mask"
                    seen_channels |= a.reshape(215, W).any(axis=1)
                    expect_rows += 1
                    checked += 1
        assert expect_rows == rows
        stride = 1 if it < 20 else 41
        v.step_random(23, stride)
        for g in range(n):
            for _ in range(stride):
                orc.orc_game_random_step(games[g], 23, 9500 + g)
    for h in games:
        orc.orc_game_free(h)
    assert checked > (3000 if mode >= 3 else 6000)
    # every block of the extended layout was exercised with non-zero content
    # (riichi sutehai, 206-214, needs an opponent's riichi: tests/test_scenarios.py::test_ext_tile_context_channels)
    for lo, hi in ((74, 78), (78, 94), (94, 98), (98, 178), (178, 189), (189, 194), (197, 206)):
        assert mode >= 3 or seen_channels[lo:hi].any(), (lo, hi)


def test_kawa_overview_vs_oracle(orc):
    """rv_vec_encode_kawa: Observation::encode_kawa_overview (4x7x34) rows of 128 hanchan at several points of the rollout."""
    import torch

    from riichienv_b200.vec_env import VecRiichiEnv

    n = 128
    v = VecRiichiEnv(n, 2, A.RULE_DEFAULT_TENHOU, seed_base=9700)
    v.reset()
    games = [orc.orc_game_new(2, 9700 + g, 0, A.RULE_DEFAULT_TENHOU, 0) for g in range(n)]
    for h in games:
        orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    out = torch.empty((n * 3, 4, 7, 34), dtype=torch.float32, device="cuda")
    idx = torch.empty((n * 3,), dtype=torch.int32, device="cuda")
    a = np.zeros(4 * 7 * 34, np.float32)
    checked = 0
    for it in range(12):
        out.fill_(-3.0)
        rows = v.encode_kawa_overview(out=out, index=idx)
        assert (out[rows:] == -3.0).all()
        h_out, h_idx = out[:rows].cpu().numpy(), idx[:rows].cpu().numpy()
        st = A.GameState()
        r = 0
        for g in range(n):
            orc.orc_game_snapshot(games[g], C.byref(st))
            if st.is_done:
                continue
            for p in range(4):
                if (st.active_mask >> p) & 1:
                    assert h_idx[r] == g * 4 + p
                    orc.orc_game_encode_kawa(games[g], a.ctypes.data_as(C.POINTER(C.c_float)))
                    assert h_out[r].tobytes() == a.tobytes(), f"iter {it} game {g} seat {p}"
                    r += 1
                    checked += 1
        assert r == rows
        v.step_random(29, 53)
        for g in range(n):
            for _ in range(53):
                orc.orc_game_random_step(games[g], 29, 9700 + g)
    for h in games:
        orc.orc_game_free(h)
    assert checked > 1000


def test_shim_observation_encode_extended(orc):
    from riichienv_b200 import RiichiEnv

    env = RiichiEnv(game_mode=0, seed=5)
    obs = env.reset()
    b = obs[0].encode_extended()
    assert len(b) == 215 * 34 * 4
    arr = np.frombuffer(b, dtype=np.float32).reshape(215, 34)
    base = np.frombuffer(obs[0].encode(), dtype=np.float32).reshape(74, 34)
    assert (arr[:74] == base).all()                       # no melds yet: channel 30 agrees too
    assert (arr[82] == 0.5).all() and (arr[189] == np.float32(14) / np.float32(34)).all()
    # standalone encoders = channel blocks of the same row (observation/python.rs:195-1270)
    f = lambda b, *shape: np.frombuffer(b, dtype=np.float32).reshape(*shape)
    assert (f(obs[0].encode_discard_history_decay(), 4, 34) == arr[74:78]).all()
    se = f(obs[0].encode_shanten_efficiency(), 4, 4)
    assert (se[1:, :3] == 0.5).all() and se[0, 0] == arr[78, 0] and se[0, 3] == 0.0
    assert f(obs[0].encode_fuuro_overview(), 4, 4, 5, 34).sum() == 0 and f(obs[0].encode_ankan_overview(), 4, 34).sum() == 0
    av = f(obs[0].encode_action_availability(), 11)
    types = {int(a.action_type) for a in obs[0].legal_actions()}
    assert av[0] == (1.0 if 5 in types else 0.0) and av[10] == 0.0          # Riichi = 5 (action.rs:55-68); no Pass on own turn
    dc = f(obs[0].encode_discard_candidates(), 5)
    assert dc[0] == np.float32(14) / np.float32(34) and 0.0 <= dc[1] + dc[2] <= 1.0
    assert f(obs[0].encode_pass_context(), 3).sum() == 0                       # nothing discarded yet
    assert f(obs[0].encode_last_tedashis(), 3, 3).sum() == 0 and f(obs[0].encode_riichi_sutehais(), 3, 3).sum() == 0
    assert f(obs[0].encode_kawa_overview(), 4, 7, 34).sum() == 0                # nothing discarded yet


def test_sequence_features_vs_oracle(orc):
    """rv_vec_encode_seq: sparse / numeric / progression / candidates of every acting seat of 128 games, observed after
    single steps and after multi-step strides (so the per-seat event deltas span several steps), bytes equal to the
    oracle; once more with caller-supplied cursors."""
    import torch

    from riichienv_b200.vec_env import VecRiichiEnv

    n, MP = 128, 96
    v = VecRiichiEnv(n, 2, A.RULE_DEFAULT_TENHOU, seed_base=7700, log_cap_words=8192)
    v.reset()
    games = [orc.orc_game_new(2, 7700 + g, 0, A.RULE_DEFAULT_TENHOU, 1) for g in range(n)]
    for h in games:
        orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    dev = dict(device="cuda")
    sparse = torch.empty((n * 3, 25), dtype=torch.uint16, **dev)
    numeric = torch.empty((n * 3, 12), dtype=torch.float32, **dev)
    prog = torch.empty((n * 3, MP, 5), dtype=torch.uint16, **dev)
    cand = torch.empty((n * 3, 64, 4), dtype=torch.uint16, **dev)
    lens = torch.empty((n * 3, 3), dtype=torch.uint16, **dev)
    idx = torch.empty((n * 3,), dtype=torch.int32, **dev)
    o_sp, o_nu, o_pr, o_ca, o_le = (np.zeros(25, np.uint16), np.zeros(12, np.float32), np.zeros(MP * 5, np.uint16),
                                    np.zeros(64 * 4, np.uint16), np.zeros(3, np.uint16))
    u16 = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint16))
    cursor = np.zeros((n, 4), np.uint32)
    checked = long_delta = 0
    st = A.GameState()
    for it in range(70):
        explicit = it % 7 == 6          # this round passes the cursors explicitly and must not advance the internal ones
        rows = v.encode_seq(sparse=sparse, numeric=numeric, prog=prog, cand=cand, lens=lens, index=idx, game_style=1,
                            start_words=cursor.copy() if explicit else None)
        h = [t[:rows].cpu().numpy() for t in (sparse, numeric, prog, cand, lens, idx)]
        row = 0
        for g in range(n):
            orc.orc_game_snapshot(games[g], C.byref(st))
            if st.is_done:
                continue
            for p in range(4):
                if (st.active_mask >> p) & 1:
                    assert h[5][row] == g * 4 + p
                    orc.orc_game_encode_seq(games[g], p, int(cursor[g, p]), st.ev_words, 1, u16(o_sp), o_nu.ctypes.data_as(C.POINTER(C.c_float)),
                                            u16(o_pr), MP, u16(o_ca), u16(o_le))
                    for k, ref in enumerate((o_sp, o_nu, o_pr, o_ca, o_le)):
                        assert h[k][row].tobytes() == ref.tobytes(), f"iter {it} game {g} seat {p} output {k}"
                    long_delta += int(o_le[1] > 8)
                    if not explicit:
                        cursor[g, p] = st.ev_words
                    row += 1
                    checked += 1
        assert row == rows
        stride = 1 if it < 35 else 11
        v.step_random(33, stride)
        for g in range(n):
            for _ in range(stride):
                orc.orc_game_random_step(games[g], 33, 7700 + g)
    for hnd in games:
        orc.orc_game_free(hnd)
    assert checked > 6000 and long_delta > 500


def test_shim_sequence_features(orc):
    """Observation.encode_seq_* of the single-env shim (the reference's method names and byte formats) against the oracle,
    driving a whole game through RiichiEnv.step with the first legal action."""
    from riichienv_b200 import RiichiEnv

    env = RiichiEnv(game_mode="4p-red-half", seed=31)
    h = orc.orc_game_new(2, 31, 0, A.RULE_DEFAULT_TENHOU, 1)
    orc.orc_game_reset(h, 0, 0, 0, 0, None, None)
    obs = env.reset()
    cursor = [0, 0, 0, 0]
    o_sp, o_nu, o_pr, o_ca, o_le = (np.zeros(25, np.uint16), np.zeros(12, np.float32), np.zeros(512 * 5, np.uint16),
                                    np.zeros(64 * 4, np.uint16), np.zeros(3, np.uint16))
    u16 = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint16))
    st = A.GameState()
    checked = 0
    for _ in range(120):
        if env.done():
            break
        orc.orc_game_snapshot(h, C.byref(st))
        acts = (A.Action * 4)()
        for p in range(4):
            acts[p].type = A.NO_ACTION
        for p, ob in obs.items():
            orc.orc_game_encode_seq(h, p, cursor[p], st.ev_words, 0, u16(o_sp), o_nu.ctypes.data_as(C.POINTER(C.c_float)), u16(o_pr), 512,
                                    u16(o_ca), u16(o_le))
            cursor[p] = st.ev_words
            assert ob.encode_seq_sparse(0) == o_sp[: o_le[0]].tobytes()
            assert ob.encode_seq_numeric() == o_nu.tobytes()
            assert ob.encode_seq_progression() == o_pr[: 5 * o_le[1]].tobytes()
            assert ob.encode_seq_candidates() == o_ca[: 4 * o_le[2]].tobytes()
            checked += 1
        chosen = {p: ob.legal_actions()[-1] for p, ob in obs.items()}
        for p, a in chosen.items():
            acts[p] = a._to_abi()
        orc.orc_game_step(h, acts)
        obs = env.step(chosen)
    orc.orc_game_free(h)
    assert checked > 100


def test_shim_observation_encode(orc):
    from riichienv_b200 import RiichiEnv

    env = RiichiEnv(game_mode=0, seed=5)
    obs = env.reset()
    b = obs[0].encode()
    assert len(b) == 74 * 34 * 4
    arr = np.frombuffer(b, dtype=np.float32).reshape(74, 34)
    assert arr[0].sum() == len({t // 4 for t in obs[0].hand})
    assert arr[30, 0] == np.float32(136 - 14 - 1) / np.float32(70.0)
    assert bytes(obs[0].mask()) == bytes(bytearray(1 if i in {a.encode() for a in obs[0].legal_actions()} else 0 for i in range(82)))
