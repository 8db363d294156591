"""The C-ABI library loads without a GPU and exports every symbol include/riichienv_b200.h declares."""
import ctypes as C
import json
import os
import re

import pytest

from riichienv_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "riichienv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from riichienv_b200._lib import lib

    L = lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"


def test_struct_sizes_match():
    from riichienv_b200._lib import lib

    L = lib()
    for i, T in enumerate((A.GameState, A.HandQuery, A.HandResult, A.Action)):
        assert L.rv_sizeof(i) == C.sizeof(T)
    import oracle

    o = oracle.load()
    for i, T in enumerate((A.GameState, A.HandQuery, A.HandResult, A.Action)):
        assert o.orc_sizeof(i) == C.sizeof(T)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device every compute entry point must fail loudly (RV_ERR_CUDA), never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from riichienv_b200._lib import Context, RvError

    with pytest.raises(RvError):
        Context(0)


def test_product_never_imports_the_checker():
    """oracle/ and tests/hostsim are test infrastructure: nothing under riichienv_b200/ (nor its C sources) may reference them"""
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "riichienv_b200")
    bad = []
    for d, _, files in os.walk(root):
        if "_build" in d or "__pycache__" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"^\s*(import|from)\s+(oracle|tests)\b", text, re.M) or re.search(r'#include\s+"[^"]*(oracle|hostsim)/', text):
                    bad.append(os.path.join(d, f))
    assert not bad, bad


def test_host_only_entry_points():
    """rv_calculate_score / rv_wall_from_seed / rv_event_to_json are pure host helpers of the ABI."""
    from riichienv_b200._lib import events_to_json, lib
    import oracle

    L, o = lib(), oracle.load()
    out, ref = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
    for han, fu, oya, tsumo, honba, np_ in [(1, 30, 0, 0, 0, 4), (3, 30, 1, 0, 2, 4), (5, 0, 0, 1, 1, 4), (13, 0, 1, 1, 0, 3), (4, 25, 0, 1, 0, 4)]:
        L.rv_calculate_score(han, fu, oya, tsumo, honba, np_, out)
        o.orc_calculate_score(han, fu, oya, tsumo, honba, np_, ref)
        assert list(out) == list(ref)
    a, b = (C.c_uint8 * 136)(), (C.c_uint8 * 136)()
    for seed in (0, 7, 2 ** 40 + 3):
        L.rv_wall_from_seed(seed, 1, 136, a)
        o.orc_wall_from_seed(seed, 1, 136, b)
        assert bytes(a) == bytes(b) and sorted(bytes(a)) == list(range(136))
    # README.md:94 / SURVEY appendix examples of the reference's JSON layout
    w0 = lambda t, n, x, y: t | (n << 8) | (x << 16) | (y << 24)
    assert events_to_json([w0(A.EV_TSUMO, 1, 0, 56)]) == ['{"actor":0,"pai":"6p","type":"tsumo"}']
    assert events_to_json([w0(A.EV_TSUMO, 1, 0, 56)], viewer=1) == ['{"actor":0,"pai":"?","type":"tsumo"}']
    assert events_to_json([w0(A.EV_DAHAI, 1, 1, 16)]) == ['{"actor":1,"pai":"5mr","tsumogiri":false,"type":"dahai"}']
    assert events_to_json([w0(A.EV_PON, 2, 2, 53), 1 | (53 - 0 + 1 << 8) | (52 << 16) | (255 << 24)]) == [
        '{"actor":2,"consumed":["5p","5pr"],"pai":"5p","target":1,"type":"pon"}']
    assert events_to_json([w0(A.EV_ANKAN, 2, 0, 0), 0 | (1 << 8) | (2 << 16) | (3 << 24)]) == [
        '{"actor":0,"consumed":["1m","1m","1m","1m"],"pai":"1m","type":"ankan"}']
    assert events_to_json([w0(A.EV_DORA, 1, 0, 80)]) == ['{"dora_marker":"3s","type":"dora"}']
    d = lambda v: v & 0xFFFFFFFF
    hora = [w0(A.EV_HORA, 10, 3, 3), 1 | (1 << 8), 72 | (0xFFFFFF << 8), 0xFF, d(-4000), d(-2000), d(-2000), 10000, 0, 0]
    assert events_to_json(hora) == [
        '{"actor":3,"deltas":[-4000,-2000,-2000,10000],"target":3,"tsumo":true,"type":"hora","ura_markers":["1s"]}']
    ry = [w0(A.EV_RYUKYOKU, 5, 0, 0), 1500, d(-1500), 1500, d(-1500)]
    assert events_to_json(ry) == ['{"deltas":[1500,-1500,1500,-1500],"reason":"exhaustive_draw","type":"ryukyoku"}']
    # the remaining event types, as the reference builds them (keys sorted, serde_json BTreeMap):
    #   reach / reach_accepted state/mod.rs:454-455, 1556-1563; chi / daiminkan 1195-1224, 1498-1517; kakan 571-581;
    #   kita state_3p/sanma.rs:47-54; start_kyoku 1785-1819 with the per-seat masking of 2109-2131; end markers 1674, 2075-2080
    assert events_to_json([w0(A.EV_START_GAME, 1, 0, 0)]) == ['{"type":"start_game"}']
    assert events_to_json([w0(A.EV_REACH, 1, 2, 0)]) == ['{"actor":2,"type":"reach"}']
    assert events_to_json([w0(A.EV_REACH_ACCEPTED, 1, 2, 0)]) == ['{"actor":2,"type":"reach_accepted"}']
    assert events_to_json([w0(A.EV_DAHAI_TSUMOGIRI, 1, 3, 135)]) == ['{"actor":3,"pai":"C","tsumogiri":true,"type":"dahai"}']
    assert events_to_json([w0(A.EV_CHI, 2, 1, 8), 0 | (12 << 8) | (16 << 16) | (255 << 24)]) == [
        '{"actor":1,"consumed":["4m","5mr"],"pai":"3m","target":0,"type":"chi"}']
    assert events_to_json([w0(A.EV_DAIMINKAN, 2, 3, 110), 1 | (108 << 8) | (109 << 16) | (111 << 24)]) == [
        '{"actor":3,"consumed":["E","E","E"],"pai":"E","target":1,"type":"daiminkan"}']
    assert events_to_json([w0(A.EV_KAKAN, 2, 0, 91), 88 | (89 << 8) | (90 << 16) | (255 << 24)]) == [
        '{"actor":0,"consumed":["5sr","5s","5s"],"pai":"5s","type":"kakan"}']
    assert events_to_json([w0(A.EV_KITA, 1, 2, 121)]) == ['{"actor":2,"pai":"N","type":"kita"}']
    assert events_to_json([w0(A.EV_END_KYOKU, 1, 0, 0), w0(A.EV_END_GAME, 1, 0, 0)]) == ['{"type":"end_kyoku"}', '{"type":"end_game"}']
    assert events_to_json([w0(A.EV_RYUKYOKU, 5, 8 + 2, 0), 4000, 4000, d(-12000), 4000]) == [
        '{"deltas":[4000,4000,-12000,4000],"reason":"Error: Illegal Action by Player 2","type":"ryukyoku"}']
    hands = [[4 * (4 * p + k) + 1 for k in range(13)] for p in range(4)]   # copy 1 of kinds 4p .. 4p+12 (no red fives)
    th = bytes(t for h in hands for t in h)
    sk = [w0(A.EV_START_KYOKU, 19, 1, 2), 3 | (52 << 8) | (1 << 16), 25000, 24000, 26000, 25000] + [
        int.from_bytes(th[i:i + 4], "little") for i in range(0, 52, 4)]
    names = lambda h: "[" + ",".join('"%d%s"' % ((t % 36) // 4 + 1, "mps"[t // 36]) for t in h) + "]"
    full = "[" + ",".join(names(h) for h in hands) + "]"
    assert events_to_json(sk) == ['{"bakaze":"S","dora_marker":"5pr","honba":3,"kyoku":3,"kyotaku":1,"oya":2,'
                                  '"scores":[25000,24000,26000,25000],"tehais":%s,"type":"start_kyoku"}' % full]
    q13 = "[" + ",".join(['"?"'] * 13) + "]"
    masked = "[" + ",".join(names(hands[p]) if p == 2 else q13 for p in range(4)) + "]"
    assert events_to_json(sk, viewer=2)[0].split('"tehais":')[1] == masked + ',"type":"start_kyoku"}'


def test_event_renderer_refuses_records_of_the_wrong_length():
    """rv_event_to_json reads 1..19 words depending on the event type: a record whose length byte disagrees with its type is
    refused (RV_ERR_INVALID), never read past its end (the renderer was run under ASan + UBSan over 2x10^6 random records)"""
    import random

    from riichienv_b200._lib import lib

    L = lib()
    out = C.create_string_buffer(4096)
    w0 = lambda t, n, x, y: t | (n << 8) | (x << 16) | (y << 24)
    for ty, good in ((A.EV_START_KYOKU, (19, 15)), (A.EV_HORA, (10, 9)), (A.EV_RYUKYOKU, (5, 4)), (A.EV_PON, (2,)), (A.EV_ANKAN, (2,))):
        for nw in range(1, 20):
            words = (C.c_uint32 * 19)(*([w0(ty, nw, 1, 40)] + [0x04030201] * 18))
            rc = L.rv_event_to_json(words, 19, -1, out, 4096)
            if nw in good or (len(good) == 1 and nw >= good[0]):
                assert rc == nw and out.value.startswith(b"{") and out.value.endswith(b"}")
            else:
                assert rc == -1, (ty, nw, rc)
    assert L.rv_event_to_json((C.c_uint32 * 1)(w0(A.EV_HORA, 10, 0, 0)), 1, -1, out, 4096) == -1     # record longer than the buffer
    rng = random.Random(1)
    for _ in range(20000):
        n = rng.randrange(1, 12)
        words = (C.c_uint32 * n)(*[rng.getrandbits(32) for _ in range(n)])
        words[0] = w0(rng.randrange(24), rng.randrange(14), rng.randrange(6), rng.randrange(256))
        rc = L.rv_event_to_json(words, n, rng.randrange(-1, 5), out, 4096)
        assert rc == -1 or 1 <= rc <= n


def test_text_log_two_renderers():
    """The oracle writes its MJAI text at event time (oracle/json.hpp, following the reference's event builders); the product
    renders binary event words on the host (csrc/json.cpp).  Full hanchan, 4P and sanma, every viewer: the texts are equal."""
    import oracle
    from riichienv_b200._lib import events_to_json

    o = oracle.load()
    for mode, n_games in ((2, 12), (5, 8), (0, 8)):
        for seed in range(n_games):
            h = o.orc_game_new(mode, seed, 0, A.RULE_DEFAULT_TENHOU if seed % 2 == 0 else A.RULE_DEFAULT_MJSOUL, 1)
            o.orc_game_reset(h, 0, 0, 0, 0, None, None)
            for _ in range(4000):
                if not o.orc_game_random_step(h, 7, seed):
                    break
            n = o.orc_game_events(h, None, 0)
            buf = (C.c_uint32 * max(n, 1))()
            o.orc_game_events(h, buf, n)
            words = list(buf[:n])
            kinds = set()
            for viewer in [-1] + list(range(3 if mode >= 3 else 4)):
                ln = o.orc_game_mjai_log(h, viewer, None, 0)
                tb = C.create_string_buffer(ln + 1)
                o.orc_game_mjai_log(h, viewer, tb, ln + 1)
                text = tb.value.decode().split("\n")
                assert text == events_to_json(words, viewer), (mode, seed, viewer)
                kinds |= {json.loads(x)["type"] for x in text}
            assert {"start_game", "start_kyoku", "tsumo", "dahai", "end_kyoku"} <= kinds
            o.orc_game_free(h)
