"""The C-ABI library loads without a GPU and exports every symbol include/riichienv_b200.h declares."""
import ctypes as C
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
