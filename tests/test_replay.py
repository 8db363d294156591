"""Replay ingestion: the MJAI reader, LogKyoku::steps' set-up and GameState::apply_log_action.

Pins, in the order of how much of the reference they carry:
  * the reference's own replay tests run unmodified through tests/test_reference_suite.py (tests/test_mjai_replay.py and
    TestReplayFuriten of tests/env/test_apply_event.py);
  * the real 12-kyoku game of the reference's fixtures (tests/data/126_204_0_mjai.jsonl, committed as the decisions of
    tests/golden/real_game_126_204_0.json plus the log itself in tests/golden/): every logged decision must be among the
    legal actions of the replayed position — 862 decisions, nine wins, calls, kans, riichi — and the round features must
    chain (`scores` of round k+1 == `end_scores` of round k);
  * logs written by this repo's own simulator (oracle, greedy agent, 4P and sanma) read back through the parser: the replayed
    record must follow the game, and oracle and kernel code (host compile here, the GPU under -m gpu) must agree on the full
    record after every log action.
"""
import ctypes as C
import gzip
import json
import os

import pytest

from riichienv_b200 import _abi as A
from tests.backends import BACKENDS, HostsimBackend, OracleBackend

HERE = os.path.dirname(os.path.abspath(__file__))
REAL_LOG = os.path.join(HERE, "golden", "126_204_0_mjai.jsonl")


def _lib():
    from riichienv_b200._lib import lib

    return lib()


def parse_text(text, rule=A.RULE_DEFAULT_TENHOU):
    """-> [(A.LogKyoku, [A.LogAction])] through the library's reader (host code: no GPU needed)"""
    from riichienv_b200._lib import check

    L = _lib()
    data = text.encode()
    h = C.c_void_p()
    check(L.rv_replay_from_text(data, len(data), rule, C.byref(h)))
    out = []
    for r in range(L.rv_replay_num_rounds(h)):
        k = A.LogKyoku()
        check(L.rv_replay_kyoku(h, r, C.byref(k)))
        acts = (A.LogAction * max(1, k.n_actions))()
        n = C.c_int(0)
        check(L.rv_replay_actions(h, r, acts, k.n_actions, C.byref(n)))
        assert n.value == k.n_actions
        out.append((k, [acts[i] for i in range(k.n_actions)]))
    L.rv_replay_free(h)
    return out


def simulated_log(mode, seed, agent_seed=0xA6E27, policy=1, max_steps=6000):
    """MJAI text log of one game played by the oracle with the greedy-win agent (wins, calls, kans, riichi, kita)"""
    b = OracleBackend(mode, seed)
    b.reset()
    for _ in range(max_steps):
        if b.get_state().is_done:
            break
        b.agent_step(policy, agent_seed, seed)
    return b.events_json(-1)


def begin(backend_cls, k, seed=7):
    b = backend_cls(3 if k.np == 3 else 0, seed, k.rule_bits)
    if isinstance(b, OracleBackend):
        b.lib.orc_game_replay_begin(b.h, C.byref(k))
        b.apply = lambda a: b.lib.orc_game_apply_log_action(b.h, C.byref(a))
    elif isinstance(b, HostsimBackend):
        b.lib.hs_game_replay_begin(b.h, C.byref(k))
        b.apply = lambda a: b.lib.hs_game_apply_log_action(b.h, C.byref(a))
    else:
        b.v.replay_begin((A.LogKyoku * 1)(k))
        b.apply = lambda a: b.v.apply_log_actions((A.LogAction * 1)(a))
    return b


# ------------------------------------------------------------------------------------------------ the reader
def test_reader_real_game_rounds_and_features():
    rounds = parse_text(open(REAL_LOG).read())
    assert len(rounds) == 12
    k0 = rounds[0][0]
    assert [k0.scores[p] for p in range(4)] == [25000] * 4
    assert [k0.end_scores[p] for p in range(4)] == [21000, 22000, 23000, 34000]       # tests/test_mjai_replay.py:96-98
    for (a, _), (b, _) in zip(rounds, rounds[1:]):
        assert [a.end_scores[p] for p in range(4)] == [b.scores[p] for p in range(4)]
    # every round: 13-tile deals, the dealer's first tile arrives as a DealTile
    for k, acts in rounds:
        assert k.np == 4 and [k.hand_len[p] for p in range(4)] == [13] * 4 and k.oya_drawn_tile == 255
        assert k.oya == k.ju % 4
        assert acts[0].type == A.LA_DEAL and acts[0].seat == k.oya
        assert acts[-1].type in (A.LA_HULE, A.LA_NOTILE)
    n_hule = sum(1 for _, acts in rounds for a in acts if a.type == A.LA_HULE)
    assert n_hule == 9


def test_reader_gzip_and_plain_files_agree(tmp_path):
    from riichienv_b200.replay import MjaiReplay

    text = open(REAL_LOG).read()
    plain, gz = tmp_path / "g.jsonl", tmp_path / "g.bin"          # gzip is detected by content, not by extension
    plain.write_text(text)
    with gzip.open(gz, "wt") as f:
        f.write(text)
    a, b = MjaiReplay.from_jsonl(str(plain)), MjaiReplay.from_jsonl(str(gz))
    assert a.num_rounds() == b.num_rounds() == 12
    for x, y in zip(a.take_kyokus(), b.take_kyokus()):
        assert x.events() == y.events() and x.grp_features() == y.grp_features()
    ev = next(iter(a.take_kyokus())).events()
    assert ev[0]["name"] == "NewRound" and ev[0]["data"]["left_tile_count"] < 70 and ev[1]["name"] == "DealTile"


def test_reader_errors_and_tile_names(tmp_path):
    from riichienv_b200.replay import MjaiReplay

    with pytest.raises(ValueError, match="Failed to open file"):
        MjaiReplay.from_jsonl(str(tmp_path / "missing.jsonl"))
    with pytest.raises(ValueError, match="Unknown rule"):
        MjaiReplay.from_jsonl(REAL_LOG, rule="nope")
    with pytest.raises(ValueError, match="Parse error"):
        MjaiReplay.from_text('{"type":"tsumo","actor":0}\n')                      # `pai` is a required field of the variant
    with pytest.raises(ValueError, match="Parse error"):
        MjaiReplay.from_text('{"type":"tsumo","actor":0,"pai":"1m"\n')
    r = MjaiReplay.from_text('{"type":"mystery","x":1}\n')                        # serde(other): unknown types are skipped
    assert r.num_rounds() == 0
    start = {"type": "start_kyoku", "bakaze": "S", "kyoku": 3, "honba": 2, "kyotaku": 1, "oya": 2, "scores": [1, 2, 3, 4],
             "dora_marker": "5sr", "tehais": [["5mr", "5m", "0p", "5p", "1z", "E", "7z", "C", "9s", "1m", "P", "F", "N"]] * 4}
    (k, acts), = parse_text(json.dumps(start) + "\n")                              # `kyotaku` is an accepted alias
    assert (k.chang, k.ju, k.ben, k.liqibang, k.doras[0]) == (1, 2, 2, 1, 88)
    assert [k.hands[0][i] for i in range(13)] == [16, 17, 52, 53, 108, 108, 132, 132, 104, 0, 124, 128, 120]
    assert k.n_actions == 0 and not acts


def test_reader_survives_malformed_input():
    """the C++ reader must refuse, not crash on, anything that is not a log (the data loaders skip unparseable files:
    riichienv-ml/src/riichienv_ml/datasets/mjai_logs.py:86-90)"""
    import random

    from riichienv_b200.replay import MjaiReplay, MjSoulReplay

    good = open(REAL_LOG).read()
    rng = random.Random(7)
    bad_inputs = ["{", "[1,2", '{"type":', '{"type":"start_kyoku"}', '{"type":"dahai","actor":"x","pai":1,"tsumogiri":0}',
                  "[" * 100 + "]" * 100, '{"type":"tsumo","actor":1e99,"pai":"1m"}', "\x00\x01\x02", '"just a string"', "nul"]
    for text in bad_inputs:
        try:
            MjaiReplay.from_text(text + "\n")
        except ValueError:
            pass
    for _ in range(60):                                   # truncations and byte flips of a real log
        cut = rng.randrange(1, len(good))
        text = good[:cut]
        if rng.random() < 0.5:
            i = rng.randrange(len(text))
            text = text[:i] + chr(rng.randrange(32, 127)) + text[i + 1:]
        try:
            r = MjaiReplay.from_text(text)
            assert 0 <= r.num_rounds() <= 12
        except ValueError:
            pass
    for obj in ({"data": 3}, {"data": [[{"name": "NewRound"}]]}, {"data": [[]]}, [[{"name": "NewRound", "data": {"scores": "x"}}]], [3]):
        try:
            MjSoulReplay.from_dict(obj)
        except ValueError:
            pass
    assert MjaiReplay.from_text(good).num_rounds() == 12   # and the reader is still sane afterwards


@pytest.mark.parametrize("source", ["real", "sim4", "sim3"])
def test_reader_flat_line_scanner_equals_dom_path(source):
    """tsumo / dahai lines are read by a scanner that builds no DOM; anything it does not recognise takes the DOM path.  The same
    log with an (ignored) nested member in every line — which the scanner refuses — with members reordered and with spaces must
    give byte-identical records"""
    text = open(REAL_LOG).read() if source == "real" else "\n".join(simulated_log(2 if source == "sim4" else 5, 37)) + "\n"
    want = [(bytes(k), [bytes(a) for a in acts]) for k, acts in parse_text(text)]
    assert sum(len(a) for _, a in want) > 500
    lines = [l for l in text.split("\n") if l.strip()]
    nested = "\n".join('{"zz":[],' + l.strip()[1:] for l in lines) + "\n"
    reordered = "\n".join(json.dumps(dict(reversed(list(json.loads(l).items()))), separators=(" , ", " : ")) for l in lines) + "\n"
    for variant in (nested, reordered):
        got = [(bytes(k), [bytes(a) for a in acts]) for k, acts in parse_text(variant)]
        assert got == want
    # what the scanner must hand over rather than guess: escapes, floats, negative / out-of-range seats, a missing member
    head = "\n".join(lines[:2]) + "\n"
    base = parse_text(head + '{"type":"tsumo","actor":1,"pai":"5m"}\n')
    for line in ('{"type":"tsumo","actor":1.0,"pai":"5m"}', '{"type":"tsu\\u006do","actor":1,"pai":"5m"}', '{"type":"tsumo","actor":1,"pai":"5\\u006d"}'):
        got = parse_text(head + line + "\n")
        assert [bytes(a) for a in got[0][1]] == [bytes(a) for a in base[0][1]], line
    seat0 = parse_text(head + '{"type":"tsumo","actor":0,"pai":"5m"}\n')
    for line in ('{"type":"tsumo","actor":-1,"pai":"5m"}', '{"type":"tsumo","actor":7,"pai":"5m"}', '{"type":"tsumo","actor":12345678901,"pai":"5m"}'):
        assert [bytes(a) for a in parse_text(head + line + "\n")[0][1]] == [bytes(a) for a in seat0[0][1]], line
    # random draw / discard lines — member order, spacing, odd seats and tile names, extra members — read by the scanner and,
    # with a nested member that the scanner refuses, by the DOM path: same records or the same refusal
    import random
    rng = random.Random(11)
    for _ in range(600):
        ty = rng.choice(["tsumo", "dahai"])
        members = [("type", json.dumps(ty)), ("actor", rng.choice(["0", "1", "2", "3", "4", "9", "007", "123456789", "1234567890"])),
                   ("pai", json.dumps(rng.choice(["5m", "5mr", "E", "C", "9s", "?", "", "0p", "5m5", "x", "１ｍ"])))]
        if ty == "dahai" or rng.random() < 0.2:
            members.append(("tsumogiri", rng.choice(["true", "false"])))
        if rng.random() < 0.3:
            members.append((rng.choice(["note", "ts", "flag"]), rng.choice(['"abc"', "17", "true", "false"])))
        if rng.random() < 0.2:
            members.pop(rng.randrange(1, len(members)))          # a member missing (possibly a required one)
        rng.shuffle(members)
        sp = lambda: rng.choice(["", "", " ", "  ", "\t"])
        line = "{" + ",".join(f"{sp()}{json.dumps(k)}{sp()}:{sp()}{v}{sp()}" for k, v in members) + "}" + sp()
        forced = '{"zz":{"a":[1]},' + line[1:]
        def read(text):
            try:
                return [[bytes(x) for x in acts] for _, acts in parse_text(text)]
            except ValueError as e:
                return str(e)
        assert read(head + line + "\n") == read(head + forced + "\n"), line
    for line, msg in (('{"type":"tsumo","actor":1}', "missing field `pai`"), ('{"type":"dahai","actor":1,"pai":"5m"}', "missing field `tsumogiri`"),
                      ('{"type":"tsumo","actor":"1","pai":"5m"}', "invalid type"), ('{"type":"tsumo","actor":1,"pai":"5m"} x', "trailing characters")):
        with pytest.raises(ValueError, match=msg):
            parse_text(head + line + "\n")


def test_reader_fuzz_under_sanitizers(tmp_path):
    """tests/fuzz/replay_fuzz.cpp: csrc/replay.cpp compiled with AddressSanitizer + UBSan, fed mutated paifu (with paishan, per-deal
    dora lists and tile counts, fans) and mutated MJAI logs through every rv_replay_* entry point — no report, and the queries
    of the win-context walk stay inside their arrays whatever the log says"""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    R = _shim("oracle")
    names = [f"{n}{s}" for s in "mps" for n in range(1, 10)] + [f"{n}z" for n in range(1, 8)]
    wall = "".join(names[(i * 5 + 3) % 34] for i in range(136))
    rounds = []
    for text in (open(REAL_LOG).read(), "\n".join(simulated_log(2, 40)) + "\n", "\n".join(simulated_log(5, 60)) + "\n"):
        for i, (_, r) in enumerate(_paifu_rounds(R, text, False)):
            if i % 2 == 0:
                r[0]["data"]["paishan"] = wall
            left = 70
            for e in r:
                if e["name"] == "DealTile":
                    left -= 1
                    if i % 3:
                        e["data"]["left_tile_count"] = max(0, left)
                    if i % 4 == 1:
                        e["data"]["doras"] = r[0]["data"]["doras"]
                if e["name"] == "Hule":
                    for h in e["data"]["hules"]:
                        h["fans"] = [{"id": 2, "val": 1}, {"id": 31, "val": 2}]
            rounds.append(r)
    paifu = tmp_path / "paifu.json"
    paifu.write_text(json.dumps({"rounds": rounds}))
    exe = tmp_path / "replay_fuzz"
    root = os.path.dirname(HERE)
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-o", str(exe),
                         os.path.join(HERE, "fuzz", "replay_fuzz.cpp"), os.path.join(root, "riichienv_b200", "csrc", "replay.cpp"), "-lz"],
                        capture_output=True, text=True, cwd=os.path.join(HERE, "fuzz"))
    if cc.returncode != 0 and "sanitize" in cc.stderr:
        pytest.skip("sanitizer runtime not available: " + cc.stderr[-200:])
    assert cc.returncode == 0, cc.stderr[-2000:]
    out = subprocess.run([str(exe), str(paifu), REAL_LOG, "400"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "runtime error" not in out.stderr and "AddressSanitizer" not in out.stderr, out.stderr[-3000:]
    ok, bad, ctxs = (int(out.stdout.split()[i]) for i in (1, 3, 6))
    assert ok > 100 and bad > 50 and ctxs > 1000, out.stdout


# ------------------------------------------------------------------------------------------------ state tracking
def _owed(s):
    return [p for p in range(4) if not s.is_done and ((s.phase == 0 and s.current_player == p) or (s.phase == 1 and (s.active_mask >> p) & 1))]


def _replay_and_compare(rounds, backends, check_follow=None):
    steps = 0
    for k, acts in rounds:
        bs = [begin(cls, k) for cls in backends]
        ref = bs[0].get_state()
        for b in bs[1:]:
            assert A.state_fields_equal(ref, b.get_state()) == [], "after replay_begin"
        for i, a in enumerate(acts):
            for b in bs:
                b.apply(a)
            ref = bs[0].get_state()
            for b in bs[1:]:
                d = A.state_fields_equal(ref, b.get_state())
                assert d == [], f"action {i} (type {a.type}) of a kyoku: {d}"
            for p in _owed(ref):
                want = bs[0].legal_tuples(p)
                for b in bs[1:]:
                    assert b.legal_tuples(p) == want
            if check_follow:
                check_follow(k, i, a, ref)
            steps += 1
    return steps


def test_real_game_kernel_code_equals_oracle():
    rounds = parse_text(open(REAL_LOG).read())
    assert _replay_and_compare(rounds, [OracleBackend, HostsimBackend]) == sum(len(a) for _, a in rounds)


@pytest.mark.parametrize("mode,seeds", [(2, range(3)), (5, range(3)), (0, range(100, 108)), (3, range(100, 108))])
def test_simulated_logs_replay_follows_the_game(mode, seeds):
    """logs of this repo's simulator read back: after every action hands, melds, rivers and riichi flags of the replayed record
    are what the log implies, on the oracle and on the kernel code alike"""
    n_kyoku = calls = kans = reaches = kitas = 0
    for seed in seeds:
        lines = simulated_log(mode, seed)
        rounds = parse_text("\n".join(lines) + "\n")
        n_kyoku += len(rounds)
        for k, acts in rounds:
            calls += sum(a.type == A.LA_CHI_PENG_GANG for a in acts)
            kans += sum(a.type == A.LA_ANGANG_ADDGANG for a in acts)
            reaches += sum(a.type == A.LA_DISCARD and a.flags & 1 for a in acts)
            kitas += sum(a.type == A.LA_BABEI for a in acts)

        def follow(k, i, a, s):
            if a.type == A.LA_DISCARD:
                assert s.river[a.seat][s.n_river[a.seat] - 1] == a.tile and s.last_discard_tile == a.tile
                assert a.tile not in [s.hand[a.seat][j] for j in range(s.hand_len[a.seat])] or True
                if a.flags & 1:
                    assert s.flags[a.seat] & A.F_RIICHI_DECLARED
            if a.type == A.LA_DEAL:
                assert s.drawn_tile == a.tile and s.current_player == a.seat and s.hand_len[a.seat] + 3 * s.n_melds[a.seat] == 14
            if a.type in (A.LA_HULE, A.LA_NOTILE):
                assert s.is_done

        _replay_and_compare(rounds, [OracleBackend, HostsimBackend], follow)
    assert n_kyoku >= len(list(seeds)) and calls > 0 and reaches > 0
    if mode in (3, 5):
        assert kitas > 0
    else:
        assert kans > 0 or mode == 0


def _shim(backend):
    from tests.test_real_game_replay import _env_module

    _env_module(backend)
    import riichienv_b200.replay as R

    return R


@pytest.mark.parametrize("backend", ["oracle", "hostsim"])
def test_real_game_every_logged_decision_is_legal(backend):
    """Kyoku.steps over the reference's real game: no desync (each logged action is in the legal list of the position it was
    taken in), the action stream re-reads as the log's decisions, masks cover the chosen ids"""
    R = _shim(backend)
    replay = R.MjaiReplay.from_jsonl(REAL_LOG)
    logged = [json.loads(l) for l in open(REAL_LOG)]
    want = {"dahai": 0, "reach": 0, "pon": 0, "chi": 0, "kan": 0, "daiminkan": 0, "ankan": 0, "kakan": 0, "hora": 0}
    for e in logged:
        if e["type"] in want:
            want[e["type"]] += 1
    got = {}
    n = passes = 0
    for kyoku in replay.take_kyokus():
        for pid, obs, act in kyoku.steps(seat=None, skip_single_action=False):
            n += 1
            name = act.action_type.name
            got[name] = got.get(name, 0) + 1
            m = obs.mask()
            assert m[act.encode()] == 1
            if name == "PASS":
                passes += 1
                assert any(a.action_type.name in ("PON", "CHI", "RON", "DAIMINKAN") for a in obs.legal_actions())
    assert got["DISCARD"] == want["dahai"] and got["RIICHI"] == want["reach"]
    assert got.get("PON", 0) == want["pon"] and got.get("CHI", 0) == want["chi"]
    assert got.get("ANKAN", 0) == want["ankan"] and got.get("KAKAN", 0) == want["kakan"]
    assert got.get("TSUMO", 0) + got.get("RON", 0) == want["hora"] == 9
    assert passes > 50 and n == sum(got.values())
    # per-seat iteration with forced decisions skipped is a subsequence of the full one
    k0 = next(iter(replay.take_kyokus()))
    full = [(p, a.action_type, a.tile) for p, _, a in k0.steps(seat=None, skip_single_action=False)]
    mine = [(2, a.action_type, a.tile) for _, a in k0.steps(2)]
    it = iter(full)
    assert all(x in it for x in mine) and 0 < len(mine) < len(full)


@pytest.mark.parametrize("backend", ["oracle", "hostsim"])
@pytest.mark.parametrize("mode", [2, 5])
def test_simulated_logs_every_logged_decision_is_legal(backend, mode):
    R = _shim(backend)
    for seed in (11, 12):
        text = "\n".join(simulated_log(mode, seed)) + "\n"
        replay = R.MjaiReplay.from_text(text, rule="tenhou")
        kinds = set()
        for kyoku in replay.take_kyokus():
            for pid, obs, act in kyoku.steps(seat=None, skip_single_action=False):
                kinds.add(act.action_type.name)
                assert obs.mask()[act.encode()] == 1
        assert {"DISCARD", "RIICHI", "PASS"} <= kinds and ("KITA" in kinds) == (mode == 5)


@pytest.mark.parametrize("source", ["real", "sim4", "sim3"])
def test_mjsoul_paifu_reader_round_trip(source, tmp_path):
    """MjSoulReplay reads the record format Kyoku.events() writes (replay/mjsoul_replay.rs RawAction <-> replay/mod.rs:1294-1500):
    a log read as MJAI, written out as paifu rounds and read back must give the same kyoku records, the same events and the same
    decisions.  (The reference writes a kan dora as "Dora" and reads it as "dora": renamed on the way.)"""
    R = _shim("oracle")
    text = open(REAL_LOG).read() if source == "real" else "\n".join(simulated_log(2 if source == "sim4" else 5, 21)) + "\n"
    a = R.MjaiReplay.from_text(text, rule="mjsoul")
    rounds = []
    for k in a.take_kyokus():
        ev = k.events()
        for e in ev:
            if e["name"] == "Dora":
                e["name"] = "dora"
        rounds.append(ev)
    b = R.MjSoulReplay.from_dict({"header": {}, "data": rounds})
    path = tmp_path / "paifu.json.gz"
    with gzip.open(path, "wt") as f:
        json.dump({"rounds": rounds}, f)
    c = R.MjSoulReplay.from_json(str(path))
    assert a.num_rounds() == b.num_rounds() == c.num_rounds() > 0
    n_steps = 0
    for x, y, z in zip(a.take_kyokus(), b.take_kyokus(), c.take_kyokus()):
        for other in (y, z):
            assert (x.scores, x.hands, x.doras, x.chang, x.ju, x.ben, x.liqibang, x.wliqi) == \
                   (other.scores, other.hands, other.doras, other.chang, other.ju, other.ben, other.liqibang, other.wliqi)
            assert x.events() == other.events()
        assert y._views[0].type == A.LA_NONE                      # the NewRound placeholder (Action::Other)
        sx = [(p, act.action_type, act.tile) for p, _, act in x.steps(None, skip_single_action=False)]
        sy = [(p, act.action_type, act.tile) for p, _, act in y.steps(None, skip_single_action=False)]
        assert sx == sy
        n_steps += len(sx)
    assert n_steps > 100
    # non-final rounds: end_scores = the next round's start scores; the final one: the last round replayed to its end
    ks = list(b.take_kyokus())
    for p, q in zip(ks, ks[1:]):
        assert p.end_scores == q.scores
    assert ks[0].game_end_scores == ks[-1].end_scores and len(ks[-1].end_scores) == len(ks[-1].scores)
    with pytest.raises(ValueError, match="missing 'data'"):
        R.MjSoulReplay.from_dict({"header": {}})
    with pytest.raises(ValueError, match="expected dict or list"):
        R.MjSoulReplay.from_dict(3)


# ------------------------------------------------------------------------------------------------ win contexts / verify
def _paifu_rounds(R, text, sanma):
    """a log of this repo's simulator turned into paifu rounds as MjSoul records them: Kyoku.events() of the MJAI reading, with the
    NewRound record holding only the first dora marker and the count of the live wall before the dealer's draw"""
    rounds = []
    for k in R.MjaiReplay.from_text(text, rule="mjsoul").take_kyokus():
        ev = k.events()
        ev[0]["data"]["doras"] = ev[0]["data"]["doras"][:1]
        ev[0]["data"]["left_tile_count"] = 55 if sanma else 70
        for e in ev:
            if e["name"] == "Dora":
                e["name"] = "dora"
        rounds.append((k, ev))
    return rounds


def _check_payments(ctx_list, k, hora_events, np_):
    """the evaluator's answer for the walk's hand + conditions against what the game paid (honba and sticks are not the
    walk's business: WinResultContextIterator evaluates with honba 0, replay/mod.rs:2008-2009)"""
    honba = k.ben
    paid = {}
    for c, e in zip(ctx_list, hora_events):
        r = c.actual
        assert r.is_win and c.seat == e["actor"], (c, e)
        if e["actor"] == e["target"]:                    # tsumo: every other seat pays its share + 100 per honba
            oya = k.ju % np_
            for p in range(np_):
                if p != c.seat:
                    share = r.tsumo_agari_oya if p == oya else r.tsumo_agari_ko
                    paid[p] = paid.get(p, 0) + share + 100 * honba
        else:                                           # ron: the discarder pays ron_agari + 300 (200 in sanma) per honba
            paid[e["target"]] = paid.get(e["target"], 0) + r.ron_agari + (np_ - 1) * 100 * honba
    return paid


@pytest.mark.parametrize("backend", ["oracle", "hostsim"])
@pytest.mark.parametrize("mode,seeds", [(2, range(40, 52)), (5, range(60, 68))])
def test_win_contexts_of_simulated_games_pay_what_the_game_paid(backend, mode, seeds):
    """WinResultContextIterator (replay/mod.rs:1594-2093): hands, melds, dora markers and win conditions tracked by the walk
    over a log, evaluated, must give the payments the game itself made — the game derives riichi / ippatsu / haitei / rinshan /
    chankan from its own state, the walk from the log alone."""
    R = _shim(backend)
    sanma = mode >= 3
    np_ = 3 if sanma else 4
    wins = flags = 0
    for seed in seeds:
        lines = simulated_log(mode, seed)
        text = "\n".join(lines) + "\n"
        horas, cur = [], None
        for l in lines:
            e = json.loads(l)
            if e["type"] == "start_kyoku":
                cur = []
                horas.append(cur)
            elif e["type"] == "hora":
                cur.append(e)
        rounds = _paifu_rounds(R, text, sanma)
        game = R.MjSoulReplay.from_dict({"header": {}, "data": [ev for _, ev in rounds]})
        assert game.num_rounds() == len(horas)
        for k, hs in zip(game.take_kyokus(), horas):
            ctxs = list(k.take_win_result_contexts())
            assert len(ctxs) == len(hs)
            if not hs:
                continue
            # pao (a liable seat pays for a yakuman) is the one case where the payers differ from the plain split
            if any(c.actual.yakuman for c in ctxs):
                continue
            paid = _check_payments(ctxs, k, hs, np_)
            deltas = [0] * np_
            for e in hs:                                 # the log's deltas are per hora event
                deltas = [a + b for a, b in zip(deltas, e["deltas"][:np_])]
            if sanma:
                # the iterator builds the 4-player evaluator also for sanma logs (replay/mod.rs:2035): no nukidora, no 1m<->9m
                # marker wrap, so only the walk is checked: who won, on what, with how many kita set aside
                for c in ctxs:
                    n_kita = sum(v.type == A.LA_BABEI and v.seat == c.seat for v in k._views)
                    assert c.conditions.kita_count == n_kita and not c.conditions.is_sanma
                    assert len(c.tiles) + 3 * len(c.melds) == 14
                wins += len(ctxs)
                flags += sum(c.conditions.kita_count > 0 for c in ctxs)
                continue
            for p, amount in paid.items():
                # a loser who declared riichi in this kyoku also lost the stick: the hora deltas do not include it
                assert -deltas[p] == amount, (seed, k.chang, k.ju, k.ben, p, deltas, paid, [c.conditions for c in ctxs])
            wins += len(ctxs)
            flags += sum(c.conditions.ippatsu or c.conditions.haitei or c.conditions.houtei or c.conditions.rinshan or
                         c.conditions.chankan or c.conditions.riichi for c in ctxs)
    assert wins > 40 and flags > 10, (wins, flags)


def test_win_contexts_read_markers_off_the_wall_and_verify_counts_mismatches():
    """a paifu with `paishan`: dora / ura markers come from the wall (replay/mod.rs:1679-1728), ankan reveals at once, the
    recorded fans / count / fu are compared by MjSoulReplay.verify (mjsoul_replay.rs:357-430)"""
    R = _shim("oracle")
    names = [f"{n}{s}" for s in "mps" for n in range(1, 10)] + [f"{n}z" for n in range(1, 8)]
    wall = [names[(i * 7) % 34] for i in range(136)]
    paishan = "".join(wall)
    # seat 0 (dealer): 1m x4 + 2m3m4m 5p6p7p 7s8s + 9s9s  -> ankan 1m, riichi, tsumo 6s / 9s
    hand0 = ["1m", "1m", "1m", "1m", "2m", "3m", "4m", "5p", "6p", "7p", "7s", "8s", "9s", "9s"]
    others = [["2p", "2p", "3p", "3p", "4p", "4p", "6m", "6m", "7m", "7m", "8m", "8m", "1z"] for _ in range(3)]
    new_round = {"scores": [25000] * 4, "doras": [wall[136 - 5]], "tiles0": hand0, "tiles1": others[0], "tiles2": others[1],
                 "tiles3": others[2], "chang": 0, "ju": 0, "ben": 0, "liqibang": 0, "left_tile_count": 69, "paishan": paishan}
    def hule(fans, count, fu):
        return {"name": "Hule", "data": {"hules": [{"seat": 0, "hu_tile": "6s", "zimo": True, "count": count, "fu": fu,
                                                      "fans": [{"id": y, "val": 1} for y in fans]}]}}
    actions = [{"name": "NewRound", "data": new_round},
               {"name": "AnGangAddGang", "data": {"seat": 0, "type": 3, "tiles": "1m"}},
               {"name": "DealTile", "data": {"seat": 0, "tile": "2z", "left_tile_count": 68}},
               {"name": "DiscardTile", "data": {"seat": 0, "tile": "2z", "is_liqi": True, "is_wliqi": True}},
               {"name": "DealTile", "data": {"seat": 1, "tile": "3z", "left_tile_count": 67}},
               {"name": "DiscardTile", "data": {"seat": 1, "tile": "3z"}},
               {"name": "DealTile", "data": {"seat": 2, "tile": "3z", "left_tile_count": 66}},
               {"name": "DiscardTile", "data": {"seat": 2, "tile": "3z"}},
               {"name": "DealTile", "data": {"seat": 3, "tile": "3z", "left_tile_count": 65}},
               {"name": "DiscardTile", "data": {"seat": 3, "tile": "3z"}},
               {"name": "DealTile", "data": {"seat": 0, "tile": "6s", "left_tile_count": 64}}]
    game = R.MjSoulReplay.from_dict({"header": {}, "data": [actions + [hule([1, 2, 30], 3, 40)]]})
    k = next(iter(game.take_kyokus()))
    assert k.paishan == paishan and k.events()[0]["data"]["paishan"] == paishan
    assert k.events()[2]["data"]["left_tile_count"] == 68
    (c,) = list(k.take_win_result_contexts())
    # the ankan revealed the second marker at once: markers = wall[-5], wall[-7]; ura = wall[-6], wall[-8]
    from riichienv_b200.replay import _tile_str
    assert [_tile_str(x) for x in c.dora_indicators] == [wall[131], wall[129]]
    assert [_tile_str(x) for x in c.ura_indicators] == [wall[130], wall[128]]
    assert c.seat == 0 and len(c.tiles) == 11 and [m.meld_type.name for m in c.melds] == ["Ankan"] and not c.melds[0].opened
    cd = c.conditions
    # the ankan ended the first go-around, so the riichi is a plain one for ippatsu purposes only in the flags the walk keeps
    assert cd.tsumo and cd.riichi and cd.double_riichi and cd.ippatsu and not cd.rinshan and not cd.haitei and not cd.tsumo_first_turn
    assert c.expected_yaku == [1, 2, 30] and (c.expected_han, c.expected_fu) == (3, 40)
    r = c.actual
    assert r.is_win and {1, 30}.issubset(r.yaku)
    assert c.calculate(c.create_calculator()).han == r.han
    # verify(): the recorded answer equal to the evaluator's is no mismatch; another yaku list, han or fu is one
    right = [y for y in r.yaku if y not in (31, 32, 33)]
    n_dora = sum(y in (31, 32, 33) for y in r.yaku)
    def counts(fans, count, fu):
        g = R.MjSoulReplay.from_dict({"header": {}, "data": [actions + [hule(fans, count, fu)]]})
        return g.verify()
    assert counts(right, r.han - n_dora, r.fu) == (1, 0)
    assert counts(right, r.han, r.fu) == (1, 0)                   # han with the dora han counted is accepted too
    assert counts(right, r.han + 1, r.fu) == (1, 1)
    assert counts(right, r.han, r.fu + 10) == (1, 1)
    assert counts(right[:-1], r.han, r.fu) == (1, 1)


def _mini_paifu(R, actions, np_=4, hands=None, paishan=None, left=69, ju=0, chang=0):
    """one round: NewRound with 13 filler tiles per seat (the dealer 14) + the given actions"""
    fill = [["1m", "2m", "3m", "4m", "5m", "6m", "7m", "8m", "9m", "1p", "2p", "3p", "4p"] for _ in range(np_)]
    fill[ju % np_] = fill[ju % np_] + ["5p"]
    hands = hands or fill
    nr = {"scores": [25000] * np_ if np_ == 4 else [35000] * 3, "doras": ["1z"], "chang": chang, "ju": ju, "ben": 0, "liqibang": 0,
          "left_tile_count": left}
    for p in range(np_):
        nr[f"tiles{p}"] = hands[p]
    if paishan:
        nr["paishan"] = paishan
        nr["doras"] = [paishan[2 * 131: 2 * 132]]
    game = R.MjSoulReplay.from_dict({"header": {}, "data": [[{"name": "NewRound", "data": nr}] + actions]})
    return list(next(iter(game.take_kyokus())).take_win_result_contexts())


def _act(name, **data):
    return {"name": name, "data": data}


def _hule(seat, tile, zimo, **kw):
    return _act("Hule", hules=[dict(seat=seat, hu_tile=tile, zimo=zimo, count=1, fu=30, fans=[], **kw)])


def test_win_context_conditions_follow_the_walk():
    """the flags WinResultContextIterator derives from the log (replay/mod.rs:1741-2060), one scenario each"""
    R = _shim("oracle")
    flags = lambda c: {n for n in ("tsumo", "riichi", "double_riichi", "ippatsu", "haitei", "houtei", "rinshan", "chankan",
                                   "tsumo_first_turn") if getattr(c.conditions, n)}
    base = ["1m", "2m", "3m", "4m", "5m", "6m", "7m", "8m", "9m", "1p", "2p"]
    H = [base + ["3p", "4p", "5p"], base + ["3p", "4p"], base + ["9p", "9p"], base + ["3p", "4p"]]   # seat 2 can pon 9p
    # tenhou: the dealer wins on the dealt hand
    (c,) = _mini_paifu(R, [_hule(0, "5p", True)])
    assert flags(c) == {"tsumo", "tsumo_first_turn"} and len(c.tiles) == 14 and c.conditions.player_wind == 0
    # riichi + ippatsu tsumo; the same with a call in between: no ippatsu, no first turn for anybody
    turn = [_act("DiscardTile", seat=0, tile="5p", is_liqi=True), _act("DealTile", seat=1, tile="9p", left_tile_count=68),
            _act("DiscardTile", seat=1, tile="9p")]
    (c,) = _mini_paifu(R, turn + [_act("DealTile", seat=0, tile="5p", left_tile_count=60), _hule(0, "5p", True)])
    assert flags(c) == {"tsumo", "riichi", "ippatsu"}
    call = [_act("ChiPengGang", seat=2, type=1, tiles=["9p", "9p", "9p"], froms=[1, 2, 2]), _act("DiscardTile", seat=2, tile="1m")]
    (c,) = _mini_paifu(R, turn + call + [_act("DealTile", seat=0, tile="5p", left_tile_count=60), _hule(0, "5p", True)], hands=H)
    assert flags(c) == {"tsumo", "riichi"}
    (c,) = _mini_paifu(R, turn + call + [_hule(1, "1m", False)], hands=H)
    assert flags(c) == set() and c.tiles[-1] == 0 and len(c.tiles) == 14            # ron: the win tile joins the hand
    (c,) = _mini_paifu(R, turn + call + [_hule(2, "1m", True)], hands=H)
    assert [m.meld_type.name for m in c.melds] == ["Pon"] and c.melds[0].from_who == 1 and c.melds[0].called_tile == 68 and c.melds[0].opened
    assert len(c.tiles) == 13 - 2 - 1 and c.conditions.player_wind == 2
    # double riichi (is_wliqi) keeps both flags
    (c,) = _mini_paifu(R, [_act("DiscardTile", seat=0, tile="5p", is_liqi=True, is_wliqi=True), _hule(0, "5p", False)])
    assert flags(c) == {"riichi", "double_riichi", "ippatsu"}
    # last tile: haitei on a draw, houtei on the discard after it — but not off a replacement tile
    (c,) = _mini_paifu(R, [_act("DiscardTile", seat=0, tile="5p"), _act("DealTile", seat=1, tile="9p", left_tile_count=0), _hule(1, "9p", True)])
    assert flags(c) == {"tsumo", "haitei", "tsumo_first_turn"}     # first turn is per seat: seat 1 has not discarded, nobody called
    (c,) = _mini_paifu(R, [_act("DiscardTile", seat=0, tile="5p"), _act("DealTile", seat=1, tile="9p", left_tile_count=0),
                           _act("DiscardTile", seat=1, tile="9p"), _hule(2, "9p", False)])
    assert flags(c) == {"houtei"}
    # ankan + replacement draw (a DealTile that carries `doras`): rinshan, and never haitei
    kan = [_act("AnGangAddGang", seat=0, type=3, tiles="1m"), _act("DealTile", seat=0, tile="5p", doras=["1z", "2z"], left_tile_count=0)]
    hands = [["1m", "1m", "1m", "1m", "5m", "6m", "7m", "8m", "9m", "1p", "2p", "3p", "4p", "5p"]] + [["2m"] * 13] * 3
    (c,) = _mini_paifu(R, kan + [_hule(0, "5p", True)], hands=hands)
    assert flags(c) == {"tsumo", "rinshan"} and [m.meld_type.name for m in c.melds] == ["Ankan"] and len(c.tiles) == 11
    assert c.dora_indicators == [108, 112]
    # kakan robbed: pon of 9p, then the added 9p is ronned by seat 3 (ippatsu of the riichi player survives until the kan resolves)
    rob = turn + call + [_act("DealTile", seat=3, tile="2z", left_tile_count=66), _act("DiscardTile", seat=3, tile="2z"),
                         _act("DealTile", seat=0, tile="2z", left_tile_count=65), _act("DiscardTile", seat=0, tile="2z"),
                         _act("DealTile", seat=1, tile="2z", left_tile_count=64), _act("DiscardTile", seat=1, tile="2z"),
                         _act("DealTile", seat=2, tile="9p", left_tile_count=63), _act("AnGangAddGang", seat=2, type=2, tiles="9p")]
    (c,) = _mini_paifu(R, rob + [_hule(3, "9p", False)], hands=H)
    assert flags(c) == {"chankan"}
    (c,) = _mini_paifu(R, rob + [_act("DealTile", seat=2, tile="3z", doras=["1z"], left_tile_count=62), _hule(2, "3z", True)], hands=H)
    assert flags(c) == {"tsumo", "rinshan"} and [(m.meld_type.name, len(m.tiles), m.from_who) for m in c.melds] == [("Kakan", 4, 1)]
    # sanma: a ron on the set-aside north uses the ippatsu of before the kita; kita are counted per seat
    h3 = [["1m", "9m", "1p", "2p", "3p", "4p", "5p", "6p", "7p", "8p", "9p", "1s", "2s", "4z"], ["1s"] * 12 + ["4z"], ["2s"] * 13]
    kita = [_act("DiscardTile", seat=0, tile="1m", is_liqi=True), _act("DealTile", seat=1, tile="4z", left_tile_count=50),
            _act("BaBei", seat=1, moqie=True)]
    (c,) = _mini_paifu(R, kita + [_hule(0, "4z", False)], np_=3, hands=h3, left=54)
    assert flags(c) == {"riichi", "ippatsu"} and c.conditions.kita_count == 0 and c.conditions.player_wind == 0
    (c,) = _mini_paifu(R, kita + [_act("DealTile", seat=1, tile="3s", doras=["1z"], left_tile_count=49), _hule(1, "3s", True)], np_=3, hands=h3, left=54)
    assert flags(c) == {"tsumo", "rinshan"} and c.conditions.kita_count == 1 and len(c.tiles) == 14   # one north set aside, two tiles drawn
    # seat winds turn with ju; the round wind is chang
    (c,) = _mini_paifu(R, [_act("DiscardTile", seat=1, tile="5p"), _hule(0, "5p", False)], ju=1, chang=1)
    assert (c.conditions.player_wind, c.conditions.round_wind) == (3, 1)


def test_win_context_kan_doras_follow_the_wall():
    """with `paishan` the markers are read off the wall: an ankan reveals at once, a daiminkan / kakan at the next discard or kan
    (replay/mod.rs:1770-1776, 1843-1852, 1905-1936)"""
    R = _shim("oracle")
    names = [f"{n}{s}" for s in "mps" for n in range(1, 10)] + [f"{n}z" for n in range(1, 8)]
    wall = [names[(i * 5 + 3) % 34] for i in range(136)]
    from riichienv_b200.replay import _tile_str
    marks = lambda c: [_tile_str(x) for x in c.dora_indicators]
    hands = [["1m"] * 14, ["2m"] * 13, ["3m"] * 13, ["4m"] * 13]
    first = [_act("DiscardTile", seat=0, tile="1m")]
    dmk = first + [_act("ChiPengGang", seat=1, type=2, tiles=["1m", "1m", "1m", "1m"], froms=[0, 1, 1, 1]),
                   _act("DealTile", seat=1, tile="7z", doras=[wall[131]], left_tile_count=60)]
    def run(actions):
        return _mini_paifu(R, actions, hands=hands, paishan="".join(wall))
    (c,) = run(dmk + [_hule(1, "7z", True)])
    assert marks(c) == [wall[131]] and c.conditions.rinshan            # the daiminkan's marker is not up at the rinshan win
    (c,) = run(dmk + [_act("DiscardTile", seat=1, tile="7z"), _hule(2, "7z", False)])
    assert marks(c) == [wall[131], wall[129]]                          # ... it is after the discard
    (c,) = run(dmk + [_act("AnGangAddGang", seat=1, type=3, tiles="2m"), _act("DealTile", seat=1, tile="6z", left_tile_count=59),
                      _hule(1, "6z", True)])
    assert marks(c) == [wall[131], wall[129], wall[127]]                # a second kan flushes the pending one; the ankan shows its own
    (c,) = run(first + [_act("DealTile", seat=1, tile="2m", left_tile_count=60), _act("DiscardTile", seat=1, tile="2m", is_liqi=True),
                        _act("DealTile", seat=2, tile="2m", left_tile_count=59), _act("DiscardTile", seat=2, tile="2m"),
                        _hule(1, "2m", False)])
    assert marks(c) == [wall[131]] and [_tile_str(x) for x in c.ura_indicators] == [wall[130]]


def _paifu_tid(t):
    """TileConverter::parse_tile_136 (replay/mod.rs:2184-2222): "0m" the red five, a plain five copy 1, anything else copy 0"""
    num, suit = int(t[0]), "mpsz".index(t[1])
    kind = 9 * suit + (5 if num == 0 else num) - 1
    return {4: 16, 13: 52, 22: 88}[kind] if num == 0 else 4 * kind + (1 if kind in (4, 13, 22) else 0)


def _contexts_as_tuples(k):
    out = []
    for c in k.take_win_result_contexts():
        q = c._c.query
        out.append((c.seat, c.tiles, [(int(m.meld_type), m.tiles) for m in c.melds], c.agari_tile, c.dora_indicators, c.ura_indicators,
                    int(q.cond), q.player_wind, q.round_wind, q.kita_count))
    return out


@pytest.mark.parametrize("mode,seeds", [(2, range(40, 48)), (5, range(60, 66))])
def test_win_context_walk_equals_the_python_restatement(mode, seeds):
    """rv_replay_win_contexts (C++, in the reader) against oracle/win_walk.py (a plain restatement of
    WinResultContextIterator::do_next): hands, melds, markers, condition bits, winds, kita — on games as the simulator played them,
    as paifu with tile counts / a wall, and with a third of every log's draws and discards dropped (hands the walk cannot
    reconcile: both must go wrong the same way)"""
    import random

    from oracle.win_walk import walk

    R = _shim("oracle")
    names = [f"{n}{s}" for s in "mps" for n in range(1, 10)] + [f"{n}z" for n in range(1, 8)]
    wall_names = [names[(i * 11 + 5) % 34] for i in range(136)]
    rng = random.Random(9)
    n_ctx = 0
    for seed in seeds:
        text = "\n".join(simulated_log(mode, seed)) + "\n"
        game = R.MjaiReplay.from_text(text)                                   # MJAI reading: no tile counts, tile_raw_id 0
        for k in game.take_kyokus():
            assert _contexts_as_tuples(k) == walk(k)
        rounds = [ev for _, ev in _paifu_rounds(R, text, mode >= 3)]
        for variant in range(3):
            rs = json.loads(json.dumps(rounds))
            for i, r in enumerate(rs):
                if variant >= 1 and i % 2 == 0:
                    r[0]["data"]["paishan"] = "".join(wall_names)
                if variant == 2:
                    r[1:] = [e for e in r[1:] if e["name"] not in ("DealTile", "DiscardTile") or rng.random() > 0.33]
            g = R.MjSoulReplay.from_dict({"header": {}, "data": rs})
            for k in g.take_kyokus():
                wall = [_paifu_tid(k.paishan[j:j + 2]) for j in range(0, len(k.paishan), 2)] if k.paishan else []
                if k._win_error:
                    continue
                got = _contexts_as_tuples(k)
                assert got == walk(k, wall), (seed, variant, k.chang, k.ju, k.ben)
                n_ctx += len(got)
    assert n_ctx > 50


def test_win_contexts_refuse_logs_the_walk_cannot_follow():
    R = _shim("oracle")
    new_round = {"scores": [35000] * 3, "doras": ["1m"], "tiles0": ["1m"] * 13, "tiles1": ["2m"] * 13, "tiles2": ["3m"] * 13,
                 "chang": 0, "ju": 0, "ben": 0, "liqibang": 0}
    bad = [{"name": "NewRound", "data": new_round}, {"name": "DealTile", "data": {"seat": 3, "tile": "1m"}}]
    game = R.MjSoulReplay.from_dict({"header": {}, "data": [bad]})
    with pytest.raises(ValueError, match="seat the kyoku does not have"):
        next(iter(game.take_kyokus())).take_win_result_contexts()
    with pytest.raises(ValueError, match="seat the kyoku does not have"):
        game.verify()


# ------------------------------------------------------------------------------------------------ sequence features of replay observations
def _progression_of_events(lines):
    """per kyoku: the (actor, type, moqie, liqi, from) tuples of the MJAI text, by the rules of process_single_event_progression
    (observation/sequence_features.rs:212-313) with the oracle's tile codes, without the start marker.  moqie as replay mode
    sees it (state/event_handler.rs:343-347): the discard is the drawn tile — by tile NAME, a log does not say which copy."""
    import oracle
    from riichienv_b200.convert import mjai_to_tid

    orc = oracle.load()
    out, cur, pending, drawn = [], None, None, {}
    for l in lines:
        e = json.loads(l)
        t = e["type"]
        if t == "start_kyoku":
            cur, pending, drawn = [], None, {}
            out.append(cur)
        elif t == "tsumo":
            drawn = {e["actor"]: e["pai"]}
        elif t == "reach":
            pending = e["actor"]
        elif t == "dahai":
            liqi = int(pending == e["actor"])
            if liqi:
                pending = None
            assert not e["tsumogiri"] or drawn.get(e["actor"]) == e["pai"]
            cur.append((e["actor"], 1 + orc.orc_seq_kan37(mjai_to_tid(e["pai"])), int(drawn.get(e["actor"]) == e["pai"]), liqi, 4))
            drawn = {}
        elif t in ("chi", "pon"):
            drawn = {}
            c = [mjai_to_tid(x) for x in e["consumed"]]
            code = (38 + orc.orc_seq_encode_chi(c[0], c[1], mjai_to_tid(e["pai"])) if t == "chi"
                    else 128 + orc.orc_seq_encode_pon(c[0], c[1], mjai_to_tid(e["pai"])))
            cur.append((e["actor"], code, 2, 2, orc.orc_seq_relative_from(e["actor"], e["target"])))
        elif t == "daiminkan":
            cur.append((e["actor"], 168 + orc.orc_seq_kan37(mjai_to_tid(e["pai"])), 2, 2, orc.orc_seq_relative_from(e["actor"], e["target"])))
        elif t == "ankan":
            cur.append((e["actor"], 205 + mjai_to_tid(e["consumed"][0]) // 4, 2, 2, 4))
        elif t == "kakan":
            cur.append((e["actor"], 239 + orc.orc_seq_kan37(mjai_to_tid(e["pai"])), 2, 2, 4))
    return out


def _check_replay_seq_features(R, seeds):
    import numpy as np

    n_obs = n_tuples = 0
    kinds = set()
    for seed in seeds:
        lines = simulated_log(2, seed)
        want = _progression_of_events(lines)
        game = R.MjaiReplay.from_text("\n".join(lines) + "\n")
        assert game.num_rounds() == len(want)
        for k, tuples in zip(game.take_kyokus(), want):
            # tuples so far after i log actions: one per discard / call / kan action
            upto = [0]
            for v in k._views:
                upto.append(upto[-1] + (v.type in (A.LA_DISCARD, A.LA_CHI_PENG_GANG, A.LA_ANGANG_ADDGANG)))
            assert upto[-1] == len(tuples)
            for pid, obs, act in k.steps(None, skip_single_action=False):
                m = upto[obs._env._n_applied]
                pr = np.frombuffer(obs.encode_seq_progression(), np.uint16).reshape(-1, 5)
                assert [tuple(r) for r in pr.tolist()] == tuples[:m], (seed, k.chang, k.ju, obs._env._n_applied)
                nu = np.frombuffer(obs.encode_seq_numeric(), np.float32)
                assert nu[0] == nu[6] == k.ben and nu[8:12].tolist() == [k.scores[(pid + i) % 4] for i in range(4)]
                sp = np.frombuffer(obs.encode_seq_sparse(), np.uint16)
                assert sp[1] == 2 + pid and sp[3] == 9 + k._k.oya
                ca = np.frombuffer(obs.encode_seq_candidates(), np.uint16).reshape(-1, 4)
                assert len(ca) <= len(obs.legal_actions())
                n_obs += 1
                n_tuples += m
            kinds |= {t[1] // 1 for t in tuples if t[1] >= 38}
    return n_obs, n_tuples, kinds


def test_replay_observations_carry_the_progression_cache():
    """encode_seq_* of replay observations (scripts/validate_logs.py:100-113 calls them on every step of a log): progression =
    the replay-mode cache, equal at every decision to the tuples of the game's own MJAI text up to there (tsumogiri flags from
    the replayed record's drawn tile, riichi from the log's flag), without the start marker"""
    n_obs, n_tuples, kinds = _check_replay_seq_features(_shim("oracle"), range(30, 33))
    assert n_obs > 1500 and n_tuples > 50_000
    assert any(38 <= t < 128 for t in kinds) and any(128 <= t < 168 for t in kinds)       # chi and pon seen


def _run_validate_logs(backend, tmp_path):
    """the reference's own log validator (scripts/validate_logs.py, unmodified; read from /root/reference or from the git-ignored
    copy __graft_entry__.build() leaves in baseline/_ref/scripts) over the real game and simulated 4P / sanma hanchan: every
    encoder on every replay observation, legality of every logged action, masks, score continuity, waits at wins"""
    import shutil
    import subprocess
    import sys

    script = next((p for p in ("/root/reference/scripts/validate_logs.py",
                               os.path.join(os.path.dirname(HERE), "baseline", "_ref", "scripts", "validate_logs.py")) if os.path.exists(p)), None)
    if script is None:
        pytest.skip("reference scripts not present")
    shutil.copy(REAL_LOG, tmp_path / "real.mjson")
    (tmp_path / "sim4.mjson").write_text("\n".join(simulated_log(2, 31)) + "\n")
    (tmp_path / "sim3.mjson").write_text("\n".join(simulated_log(5, 61)) + "\n")
    root = os.path.dirname(HERE)
    env = dict(os.environ, RV_REFSUITE_BACKEND=backend, PYTHONPATH=os.pathsep.join([os.path.join(HERE, "refsuite"), root]))
    out = subprocess.run([sys.executable, os.path.join(HERE, "refsuite", "run_validate_logs.py"), script, str(tmp_path)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "Done: 3/3 passed" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("backend", ["oracle", "hostsim"])
def test_reference_validate_logs_script_passes(backend, tmp_path):
    _run_validate_logs(backend, tmp_path)


# ------------------------------------------------------------------------------------------------ labels of the batched walk
_OWN_TURN = ("DISCARD", "RIICHI", "ANKAN", "KAKAN", "TSUMO", "KITA")


def _iterator_own_turn_samples(k):
    """{(log action index, seat): action id} of the own-turn decisions Kyoku.steps yields (the first one per log action: a
    riichi discard yields Riichi and then the discard)"""
    out = {}
    for pid, obs, act in k.steps(None, skip_single_action=False):
        if act.action_type.name in _OWN_TURN:
            out.setdefault((obs._env._n_applied, pid), (act.encode(), obs))
    return out


@pytest.mark.parametrize("source", ["real", "sim4", "sim3"])
def test_batch_labels_are_the_iterators_own_turn_decisions(source):
    """ReplayBatch.own_turn_label reads off the log, per action, what Kyoku.steps yields for the seat on turn: same positions,
    same seats, same action ids — nothing missing, nothing extra"""
    R = _shim("oracle")
    text = open(REAL_LOG).read() if source == "real" else "\n".join(simulated_log(2 if source == "sim4" else 5, 33)) + "\n"
    n = 0
    ids = set()
    for k in R.MjaiReplay.from_text(text).take_kyokus():
        np_ = k._k.np
        want = {key: v[0] for key, v in _iterator_own_turn_samples(k).items()}
        got = {}
        for t, a in enumerate(k._views):
            lab = R.ReplayBatch.own_turn_label(a, np_)
            if lab is not None:
                got[(t, lab[0])] = lab[1]
        assert got == want, (k.chang, k.ju, k.ben)
        n += len(got)
        ids |= set(got.values())
    assert n > 200 and (37 in ids or 27 in ids) and (79 in ids or 56 in ids)          # riichi and tsumo seen
    assert (59 in ids) == (source == "sim3")                                          # kita


def _log_files(tmp_path):
    """real game (plain), a 4P and a sanma simulated hanchan (gzip), one corrupt file, one missing path"""
    paths = [tmp_path / "a_real.jsonl", tmp_path / "b_sim4.jsonl.gz", tmp_path / "c_bad.jsonl", tmp_path / "d_sim3.jsonl.gz",
             tmp_path / "e_missing.jsonl", tmp_path / "f_sim4.jsonl"]
    paths[0].write_text(open(REAL_LOG).read())
    with gzip.open(paths[1], "wt") as f:
        f.write("\n".join(simulated_log(2, 34)) + "\n")
    paths[2].write_text('{"type":"start_game"}\n{"type":"start_kyoku","bakaze":"E"\n')
    with gzip.open(paths[3], "wt") as f:
        f.write("\n".join(simulated_log(5, 35)) + "\n")
    paths[5].write_text("\n".join(simulated_log(2, 36)) + "\n")
    return [str(p) for p in paths]


def test_bulk_reader_equals_file_by_file(tmp_path):
    """rv_replay_from_files (a pool of host threads) + rv_replay_flatten: the rounds of the readable files in the order of the
    paths, byte for byte what reading each file alone gives, whatever the thread count; unreadable files are counted;
    rv_replay_own_turn_labels equals the shim's own_turn_label on every action"""
    import numpy as np

    from riichienv_b200._lib import check

    R = _shim("oracle")
    L = _lib()
    paths = _log_files(tmp_path)
    single = {}
    for np_ in (4, 3):
        single[np_] = [(bytes(k), [bytes(a) for a in acts]) for p in paths if "bad" not in p and "missing" not in p
                       for k, acts in parse_text(gzip.open(p, "rt").read() if p.endswith(".gz") else open(p).read()) if k.np == np_]
    for threads in (1, 4, 0):
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        h, failed = C.c_void_p(), C.c_int(0)
        check(L.rv_replay_from_files(arr, len(paths), 0, A.RULE_DEFAULT_TENHOU, threads, C.byref(h), C.byref(failed)))
        assert failed.value == 2
        assert L.rv_replay_num_rounds(h) == len(single[4]) + len(single[3])
        for np_ in (4, 3):
            nr, na = C.c_int64(0), C.c_int64(0)
            check(L.rv_replay_totals(h, np_, C.byref(nr), C.byref(na)))
            assert nr.value == len(single[np_]) and na.value == sum(len(a) for _, a in single[np_])
            ky, acts = (A.LogKyoku * nr.value)(), (A.LogAction * na.value)()
            first, ridx = (C.c_int64 * (nr.value + 1))(), (C.c_int32 * nr.value)()
            check(L.rv_replay_flatten(h, np_, ky, acts, first, ridx))
            assert list(ridx) == sorted(ridx) and first[nr.value] == na.value
            for i, (kb, ab) in enumerate(single[np_]):
                assert bytes(ky[i]) == kb and [bytes(acts[j]) for j in range(first[i], first[i + 1])] == ab
            seat, aid = np.zeros(na.value, np.int16), np.zeros(na.value, np.int16)
            p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int16))
            check(L.rv_replay_own_turn_labels(acts, na.value, np_, p16(seat), p16(aid)))
            from riichienv_b200.replay import _ActionView
            n_lab = 0
            for j in range(na.value):
                lab = R.ReplayBatch.own_turn_label(_ActionView(acts[j]), np_)
                assert (int(seat[j]), int(aid[j])) == (lab if lab is not None else (-1, -1)), j
                n_lab += lab is not None
            assert n_lab > 200
        L.rv_replay_free(h)
    assert L.rv_replay_from_files(None, 0, 0, 0, 1, C.byref(h), None) == -1


# ------------------------------------------------------------------------------------------------ the product (GPU)
@pytest.mark.gpu
def test_gpu_replay_batch_equals_oracle():
    """rv_vec_replay_begin / rv_vec_apply_log_actions on a vector of kyoku replayed in lock-step, record by record against the
    oracle after every action (real game + simulated 4P logs; sanma in its own vector)"""
    from riichienv_b200.replay import LogKyoku, ReplayBatch
    from riichienv_b200.env import GameRule

    for mode, texts in ((2, [open(REAL_LOG).read()] + ["\n".join(simulated_log(2, s)) + "\n" for s in range(4)]),
                        (5, ["\n".join(simulated_log(5, s)) + "\n" for s in range(4)])):
        rounds = [r for t in texts for r in parse_text(t)]
        kyokus = [LogKyoku(k, (A.LogAction * len(acts))(*acts), GameRule.default_tenhou()) for k, acts in rounds]
        batch = ReplayBatch(kyokus)
        oracles = [begin(OracleBackend, k, seed=i) for i, (k, _) in enumerate(rounds)]      # ReplayBatch seeds game i with i
        for i, o in enumerate(oracles):
            assert A.state_fields_equal(o.get_state(), batch.vec.get_state(i)) == [], "after replay_begin"
        step = 0
        while batch.advance():
            for i, (k, acts) in enumerate(rounds):
                if step < len(acts):
                    oracles[i].apply(acts[step])
                d = A.state_fields_equal(oracles[i].get_state(), batch.vec.get_state(i))
                assert d == [], f"mode {mode} kyoku {i} action {step}: {d}"
            step += 1
        assert step == max(len(a) for _, a in rounds)


@pytest.mark.gpu
def test_gpu_shim_real_game_steps():
    R = _shim("gpu")
    replay = R.MjaiReplay.from_jsonl(REAL_LOG)
    n = 0
    for kyoku in list(replay.take_kyokus())[:3]:
        for pid, obs, act in kyoku.steps(seat=None, skip_single_action=False):
            n += 1
            assert obs.mask()[act.encode()] == 1
            if n % 40 == 0:
                assert len(obs.encode()) == 74 * 34 * 4          # the tensor of a by-value replay observation (device encoder)
    assert n > 150


def _verify_and_collect(R, mode, seeds):
    """games whose paifu records the evaluator's own answer for even rounds and a wrong fu for odd ones: verify() must count
    exactly the odd ones; returns the walk's queries"""
    queries = []
    for seed in seeds:
        text = "\n".join(simulated_log(mode, seed)) + "\n"
        rounds = [ev for _, ev in _paifu_rounds(R, text, mode >= 3)]
        n_bad = 0
        game = R.MjSoulReplay.from_dict({"header": {}, "data": rounds})
        for i, k in enumerate(game.take_kyokus()):
            for c in k.take_win_result_contexts():
                r = c.actual
                for h in rounds[i][-1]["data"]["hules"]:
                    if h["seat"] == c.seat:
                        h["fans"] = [{"id": y, "val": 1} for y in r.yaku]
                        h["count"], h["fu"] = r.han, r.fu + (10 if i % 2 else 0)
                n_bad += (i % 2 == 1 and r.han < 13)
                queries.append(c._c.query)
        game = R.MjSoulReplay.from_dict({"header": {}, "data": rounds})
        assert game.verify() == (sum(len(k._win_ctx) for k in game.rounds), n_bad), seed
    return queries


def test_verify_files_sums_over_games(tmp_path, capsys):
    """MjSoulReplay.verify_files: several paifu files (one unreadable) through the bulk reader, one evaluation batch"""
    R = _shim("oracle")
    paths, want_total, want_bad = [], 0, 0
    for i, seed in enumerate((40, 41, 42)):
        rounds = [ev for _, ev in _paifu_rounds(R, "\n".join(simulated_log(2, seed)) + "\n", False)]
        game = R.MjSoulReplay.from_dict({"header": {}, "data": rounds})
        for j, k in enumerate(game.take_kyokus()):
            for c in k.take_win_result_contexts():
                r = c.actual
                for h in rounds[j][-1]["data"]["hules"]:
                    if h["seat"] == c.seat:
                        h["fans"] = [{"id": y, "val": 1} for y in r.yaku]
                        h["count"], h["fu"] = r.han, r.fu + (10 if (i + j) % 3 == 0 else 0)
                want_total += 1
                want_bad += ((i + j) % 3 == 0 and r.han < 13)
        p = tmp_path / f"game{i}.json.gz"
        with gzip.open(p, "wt") as f:
            json.dump({"rounds": rounds}, f)
        paths.append(str(p))
    (tmp_path / "broken.json.gz").write_bytes(b"not gzip, not json")
    paths.insert(1, str(tmp_path / "broken.json.gz"))
    assert R.MjSoulReplay.verify_files(paths, threads=2) == (want_total, want_bad, 1)
    assert want_total > 10 and 0 < want_bad < want_total
    assert capsys.readouterr().out.count("Mismatch: seat=") == want_bad


def test_batched_round_features_equal_take_grp_features():
    """ReplayBatch.round_features_of over an rv_log_kyoku array = Kyoku.take_grp_features() of every round (replay/mod.rs:1524-1590),
    incl. rank ties (the lower seat first) and final_ranks of a paifu's game_end_scores"""
    import numpy as np

    R = _shim("oracle")
    games = [R.MjaiReplay.from_text(open(REAL_LOG).read()), R.MjaiReplay.from_text("\n".join(simulated_log(5, 62)) + "\n")]
    text = "\n".join(simulated_log(2, 46)) + "\n"
    games.append(R.MjSoulReplay.from_dict({"header": {}, "data": [ev for _, ev in _paifu_rounds(R, text, False)]}))
    for g in games:
        ks = list(g.take_kyokus())
        ks[0]._k.scores[1] = ks[0]._k.scores[0]        # a tie in the start scores
        ks[0].scores[1] = ks[0].scores[0]
        np_ = ks[0]._k.np
        got = R.ReplayBatch.round_features_of((A.LogKyoku * len(ks))(*[k._k for k in ks]), np_)
        for i, k in enumerate(ks):
            want = k.take_grp_features()
            for key in ("chang", "ju", "ben", "liqibang"):
                assert int(got[key][i]) == want[key]
            for key in ("round_initial_scores", "round_end_scores", "round_delta_scores", "round_initial_ranks", "round_end_ranks",
                        "round_delta_ranks", "final_ranks"):
                assert got[key][i].tolist() == list(want[key]), (key, i)
        assert got["round_end_scores"].shape == (len(ks), np_)


def test_bulk_reader_reads_paifu_files_too(tmp_path):
    """format 1 of rv_replay_from_files = rv_replay_from_mjsoul_json per file, rounds in path order"""
    from riichienv_b200._lib import check

    R = _shim("oracle")
    L = _lib()
    paths, want = [], []
    for i, (mode, seed) in enumerate(((2, 44), (5, 64), (2, 45))):
        rounds = [ev for _, ev in _paifu_rounds(R, "\n".join(simulated_log(mode, seed)) + "\n", mode >= 3)]
        p = tmp_path / f"p{i}.json.gz"
        with gzip.open(p, "wt") as f:
            json.dump({"rounds": rounds}, f)
        paths.append(str(p))
        h = C.c_void_p()
        check(L.rv_replay_from_mjsoul_json(str(p).encode(), A.RULE_DEFAULT_MJSOUL, C.byref(h)))
        for r in range(L.rv_replay_num_rounds(h)):
            k = A.LogKyoku()
            check(L.rv_replay_kyoku(h, r, C.byref(k)))
            want.append(bytes(k))
        L.rv_replay_free(h)
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    h, failed = C.c_void_p(), C.c_int(0)
    check(L.rv_replay_from_files(arr, len(paths), 1, A.RULE_DEFAULT_MJSOUL, 3, C.byref(h), C.byref(failed)))
    got = []
    for r in range(L.rv_replay_num_rounds(h)):
        k = A.LogKyoku()
        check(L.rv_replay_kyoku(h, r, C.byref(k)))
        got.append(bytes(k))
    n = C.c_int(0)
    check(L.rv_replay_win_contexts(h, -1, None, 0, C.byref(n)))
    L.rv_replay_free(h)
    assert failed.value == 0 and got == want and len(got) > 20 and n.value > 10
    assert {bytes(k)[0] for k in map(bytes, got)} == {3, 4}                 # both variants, in the order of the paths


def test_verify_counts_exactly_the_altered_rounds(capsys):
    assert len(_verify_and_collect(_shim("oracle"), 2, range(40, 43))) > 10
    assert "Mismatch: seat=" in capsys.readouterr().out         # the reference prints every mismatch (mjsoul_replay.rs:423-432)


@pytest.mark.gpu
def test_gpu_win_contexts_and_verify_equal_oracle():
    """MjSoulReplay.verify / WinResultContext.actual on the device: every win of 4P and sanma games (the walk's queries) evaluated
    by rv_hand_eval_batch, byte for byte the oracle's answer for the same queries"""
    import oracle
    import riichienv_b200.replay as R
    from riichienv_b200 import hand as H

    for mode, seeds in ((2, range(40, 48)), (5, range(60, 64))):
        queries = _verify_and_collect(R, mode, seeds)
        n = len(queries)
        assert n > 20
        arr = (A.HandQuery * n)(*queries)
        want = (A.HandResult * n)()
        oracle.load().orc_hand_eval(arr, want, n)
        assert bytes(H.eval_queries(queries)) == bytes(want)


@pytest.mark.gpu
def test_gpu_replay_observations_carry_the_progression_cache():
    import riichienv_b200.replay as R

    n_obs, _, _ = _check_replay_seq_features(R, [30])
    assert n_obs > 400


@pytest.mark.gpu
def test_gpu_reference_validate_logs_script_passes(tmp_path):
    _run_validate_logs("gpu", tmp_path)


@pytest.mark.gpu
def test_gpu_batch_rows_carry_their_labels():
    """the batched extraction (ReplayBatch + rv_vec_encode + labels_of_rows): every own-turn decision of every kyoku gets a row
    at its position, the logged action is legal in that row's mask, and tensors / masks equal what Kyoku.steps yields"""
    import numpy as np
    import torch
    import riichienv_b200.replay as R

    for mode, texts in ((2, [open(REAL_LOG).read(), "\n".join(simulated_log(2, 33)) + "\n"]), (5, ["\n".join(simulated_log(5, 33)) + "\n"])):
        kyokus = [k for t in texts for k in R.MjaiReplay.from_text(t).take_kyokus()]
        W, M = (27, 60) if mode == 5 else (34, 82)
        batch = R.ReplayBatch(kyokus)
        K = batch.n
        obs = torch.empty((4 * K, 74, W), dtype=torch.float32, device="cuda")
        mask = torch.empty((4 * K, M), dtype=torch.uint8, device="cuda")
        idx = torch.empty((4 * K,), dtype=torch.int32, device="cuda")
        seat, aid = batch.labels()
        want_rows = int((aid >= 0).sum())
        check_k = {0: _iterator_own_turn_samples(kyokus[0]), K - 1: _iterator_own_turn_samples(kyokus[K - 1])}
        got_rows = compared = 0
        while True:
            n = batch.vec.encode(obs=obs, mask=mask, index=idx)
            lab = batch.labels_of_rows(idx, n)
            rows = torch.nonzero(lab >= 0).flatten()
            got_rows += int(rows.numel())
            if rows.numel():
                assert bool((mask[rows, lab[rows]] == 1).all()), f"a logged action is not legal in its row's mask (position {batch.position})"
                for r in rows.tolist():
                    g, s = int(idx[r]) // 4, int(idx[r]) % 4
                    assert seat[g, batch.position] == s
                    if g in check_k:
                        a_id, o = check_k[g][(batch.position, s)]
                        assert a_id == int(lab[r])
                        assert obs[r].cpu().numpy().tobytes() == o.encode() and mask[r].cpu().numpy().tobytes() == o.mask()
                        compared += 1
            if not batch.advance():
                break
        assert got_rows == want_rows > 200, (got_rows, want_rows)
        assert compared > 50


@pytest.mark.gpu
def test_gpu_batch_from_files_equals_batch_of_kyokus(tmp_path):
    """ReplayBatch.from_files (parsed, flattened and labelled in the library) against ReplayBatch(kyokus) built round by round in
    the shim: same labels, same tensor rows / masks / row index at every position"""
    import numpy as np
    import torch
    import riichienv_b200.replay as R

    paths = _log_files(tmp_path)
    for sanma in (False, True):
        a = R.ReplayBatch.from_files(paths, sanma=sanma, threads=3)
        assert a.n_failed == 2
        kyokus = []
        for p in paths:
            try:
                kyokus += [k for k in R.MjaiReplay.from_jsonl(p).take_kyokus() if (k._k.np == 3) == sanma]
            except ValueError:
                pass
        b = R.ReplayBatch(kyokus)
        assert a.n == b.n > 5
        for x, y in zip(a.labels(), b.labels()):
            assert np.array_equal(x, y)
        W, M = (27, 60) if sanma else (34, 82)
        bufs = [(torch.zeros((4 * a.n, 74, W), device="cuda"), torch.zeros((4 * a.n, M), dtype=torch.uint8, device="cuda"),
                 torch.zeros((4 * a.n,), dtype=torch.int32, device="cuda")) for _ in range(2)]
        rows = 0
        while True:
            na = a.vec.encode(obs=bufs[0][0], mask=bufs[0][1], index=bufs[0][2])
            nb = b.vec.encode(obs=bufs[1][0], mask=bufs[1][1], index=bufs[1][2])
            assert na == nb
            for t0, t1 in zip(bufs[0], bufs[1]):
                assert torch.equal(t0[:na], t1[:na])
            assert torch.equal(a.labels_of_rows(bufs[0][2], na), b.labels_of_rows(bufs[1][2], nb))
            rows += na
            more_a, more_b = a.advance(), b.advance()
            assert more_a == more_b
            if not more_a:
                break
        assert rows > 300
