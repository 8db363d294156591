"""Batched hand evaluation front-end (HandEvaluator / Conditions / calculate_score / calculate_shanten).

Mirrors hand_evaluator.rs:24-213, types.rs:192-295, score.rs:13-52, shanten.rs:250-261; the arithmetic runs in
`hand_eval_kernel` on the GPU (rv_hand_eval_batch)."""
import ctypes as C
from dataclasses import dataclass, field

from . import _abi as A
from ._lib import Context, check, lib


@dataclass
class Conditions:  # types.rs:192-233
    tsumo: bool = False
    riichi: bool = False
    double_riichi: bool = False
    ippatsu: bool = False
    haitei: bool = False
    houtei: bool = False
    rinshan: bool = False
    chankan: bool = False
    tsumo_first_turn: bool = False
    player_wind: int = 0
    round_wind: int = 0
    riichi_sticks: int = 0
    honba: int = 0
    kita_count: int = 0
    is_sanma: bool = False
    num_players: int = 4

    def bits(self):
        names = ["tsumo", "riichi", "double_riichi", "ippatsu", "haitei", "houtei", "rinshan", "chankan", "tsumo_first_turn"]
        return sum((1 << i) for i, n in enumerate(names) if getattr(self, n))


@dataclass
class WinResult:  # types.rs:281-295
    is_win: bool = False
    yakuman: bool = False
    ron_agari: int = 0
    tsumo_agari_oya: int = 0
    tsumo_agari_ko: int = 0
    yaku: list = field(default_factory=list)
    han: int = 0
    fu: int = 0
    pao_payer: object = None
    has_win_shape: bool = False


def make_query(tiles_136, melds, win_tile, dora, ura, cond: Conditions) -> A.HandQuery:
    q = A.HandQuery()
    tiles = list(tiles_136)[:14]
    for i in range(14):
        q.tiles[i] = tiles[i] if i < len(tiles) else 255
    q.n_tiles = len(tiles)
    q.n_melds = len(melds)
    for m in range(4):
        for k in range(4):
            q.meld_tiles[m][k] = 255
    for m, md in enumerate(list(melds)[:4]):
        q.meld_type[m] = int(md.meld_type)
        for k, t in enumerate(md.tiles[:4]):
            q.meld_tiles[m][k] = t
    q.win_tile = win_tile
    q.n_dora, q.n_ura = min(5, len(dora)), min(5, len(ura))
    for i, t in enumerate(list(dora)[:5]):
        q.dora_ind[i] = t
    for i, t in enumerate(list(ura)[:5]):
        q.ura_ind[i] = t
    q.cond = cond.bits()
    q.player_wind, q.round_wind, q.honba = int(cond.player_wind), int(cond.round_wind), int(cond.honba)
    q.sanma, q.kita_count = int(bool(cond.is_sanma)), int(cond.kita_count)
    return q


def eval_queries(queries, device=0):
    n = len(queries)
    arr = (A.HandQuery * n)(*queries)
    out = (A.HandResult * n)()
    check(lib().rv_hand_eval_batch(Context.get(device).handle, arr, out, n))
    return out


def _to_result(r: A.HandResult) -> WinResult:
    return WinResult(bool(r.is_win), bool(r.yakuman), r.ron_agari, r.tsumo_agari_oya, r.tsumo_agari_ko,
                     [b for b in range(64) if (r.yaku_mask >> b) & 1], r.han, r.fu, None, bool(r.has_win_shape))


class HandEvaluator:
    def __init__(self, tiles_136, melds=()):
        self.tiles_136 = list(tiles_136)
        self.melds = list(melds)

    def calc(self, win_tile, dora_indicators=(), ura_indicators=(), conditions=None) -> WinResult:
        q = make_query(self.tiles_136, self.melds, win_tile, dora_indicators, ura_indicators, conditions or Conditions())
        return _to_result(eval_queries([q])[0])

    def _waits_mask(self):
        q = make_query(self.tiles_136, self.melds, 0, (), (), Conditions())
        return eval_queries([q])[0].wait_mask if len(self.tiles_136) + 3 * len(self.melds) == 13 else 0

    def get_waits(self):
        m = self._waits_mask()
        return [k for k in range(34) if (m >> k) & 1]

    get_waits_u8 = get_waits

    def is_tenpai(self):
        return self._waits_mask() != 0


def calculate_score(han, fu, is_oya, is_tsumo, honba=0, num_players=4):
    out = (C.c_uint32 * 4)()
    check(lib().rv_calculate_score(int(han), int(fu), int(is_oya), int(is_tsumo), int(honba), int(num_players), out))

    class Score:
        pay_ron, pay_tsumo_oya, pay_tsumo_ko, total = out[0], out[1], out[2], out[3]

    return Score


def calculate_shanten(hand_tiles) -> int:  # shanten.rs:250-261 (len_div3 = n_tiles // 3)
    tiles = [t for t in hand_tiles if t // 4 < 34][:14]
    if not tiles:
        return 0
    q = make_query(tiles, (), tiles[-1], (), (), Conditions())
    r = eval_queries([q])[0]
    # the kernel adds the win tile to 13-tile hands (HandEvaluator::calc semantics); `shanten13` is the raw 13-tile figure
    return r.shanten13 if len(tiles) == 13 else r.shanten


def calculate_shanten_3p(hand_tiles) -> int:  # shanten.rs:470-484 (1m / 9m koutsu-only, chiitoi without 2m-8m)
    tiles = [t for t in hand_tiles if t // 4 < 34][:14]
    if not tiles:
        return 0
    q = make_query(tiles, (), tiles[-1], (), (), Conditions(is_sanma=True))
    r = eval_queries([q])[0]
    return r.shanten13 if len(tiles) == 13 else r.shanten
