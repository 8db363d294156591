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

    def yaku_list(self):  # types.rs:356-361
        from .yaku_table import get_yaku_by_id

        return [y for y in map(get_yaku_by_id, self.yaku) if y is not None]


def make_query(tiles_136, melds, win_tile, dora, ura, cond: Conditions) -> A.HandQuery:
    q = A.HandQuery()
    tiles = [int(t) for t in tiles_136]
    melds = list(melds)
    # the kernel indexes its tables with these values: reject what is not a hand instead of computing garbage
    if len(tiles) > 14:
        raise ValueError(f"a hand holds at most 14 concealed tiles, got {len(tiles)}")
    if len(melds) > 4:
        raise ValueError(f"a hand holds at most 4 melds, got {len(melds)}")
    if any(not 0 <= t < 136 for t in tiles) or not 0 <= int(win_tile) < 136:
        raise ValueError("tile ids must be in 0..135")
    if any(not 0 <= int(t) < 136 for m in melds for t in m.tiles) or any(not 3 <= len(m.tiles) <= 4 for m in melds):
        raise ValueError("melds hold 3 or 4 tile ids in 0..135")
    if any(not 0 <= int(t) < 136 for t in list(dora) + list(ura)):
        raise ValueError("indicator tile ids must be in 0..135")
    for i in range(14):
        q.tiles[i] = tiles[i] if i < len(tiles) else 255
    q.n_tiles = len(tiles)
    q.n_melds = len(melds)
    for m in range(4):
        for k in range(4):
            q.meld_tiles[m][k] = 255
    for m, md in enumerate(melds):
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


# WinResult.yaku lists the ids in the order calculate_yaku pushes them, which is a fixed sequence of tests
# (yaku.rs:236-296 kokushi / chiitoitsu, 322-548 standard hands, 843-890 static yaku, 892-1055 yakuman; sanma adds
# nukidora after ura, yaku_3p.rs:706-709), so the order follows from the id set the kernel returns.
_YAKUMAN_ORDER = [39, 41, 40, 44, 47, 45, 35, 36, 48, 38, 37, 50, 43]
_STATIC_ORDER = [2, 18, 30, 1, 5, 6, 4, 3, 31, 32, 33, 34]
_ORDER_STANDARD = [49, 42] + _YAKUMAN_ORDER + _STATIC_ORDER + [12, 14, 7, 8, 9, 11, 10, 23, 21, 22, 20, 28, 13, 16, 17, 19,
                                                              29, 27, 24, 26, 15]
_ORDER_CHIITOI = [25, 12, 29, 27, 24] + _YAKUMAN_ORDER + _STATIC_ORDER


def ordered_yaku(mask: int):
    order = _ORDER_CHIITOI if (mask >> 25) & 1 else _ORDER_STANDARD
    return [y for y in order if (mask >> y) & 1]


def _to_result(r: A.HandResult) -> WinResult:
    return WinResult(bool(r.is_win), bool(r.yakuman), r.ron_agari, r.tsumo_agari_oya, r.tsumo_agari_ko,
                     ordered_yaku(r.yaku_mask), r.han, r.fu, None, bool(r.has_win_shape))


class HandEvaluator:
    """src/riichienv/hand.py:37-285 over hand_evaluator.rs:24-213 (the arithmetic runs in hand_eval_kernel)."""

    _sanma = False

    def __init__(self, tiles, melds=None):
        if isinstance(tiles, tuple) and len(tiles) == 2 and isinstance(tiles[0], list):  # parse_hand() result
            tiles, melds = tiles[0], (melds or tiles[1])
        self.tiles_136 = list(tiles)
        self.melds = list(melds or [])

    @staticmethod
    def hand_from_text(text: str) -> "HandEvaluator":  # hand.py:45-66: 13 tiles + one more per kan
        from .convert import parse_hand

        tiles, melds = parse_hand(text)
        kans = sum(1 for m in melds if int(m.meld_type) >= 2)
        have = len(tiles) + sum(len(m.tiles) for m in melds)
        if have != 13 + kans:
            raise ValueError(f"Hand must have {13 + kans} tiles (got {have})")
        return HandEvaluator(sorted(tiles), melds)

    @staticmethod
    def calc_from_text(text: str, dora_indicators=None, conditions=None, ura_indicators=None) -> WinResult:  # hand.py:95-130
        from .convert import parse_hand

        tiles, melds = parse_hand(text)
        if not tiles and not melds:
            raise ValueError("Empty hand")
        if not tiles:
            raise ValueError("No standing tiles to check for win tile")
        dora = sorted(parse_hand(dora_indicators)[0]) if dora_indicators else []
        ura = sorted(parse_hand(ura_indicators)[0]) if ura_indicators else []
        return HandEvaluator(sorted(tiles), melds).calc(tiles[-1], dora, conditions, ura)

    def to_text(self) -> str:  # hand.py:68-93: concealed tiles grouped by suit, then the melds with a 0 call index
        def digit(t):
            return "0" if t in (16, 52, 88) else str((t // 4) % 9 + 1 if t < 108 else (t - 108) // 4 + 1)

        def suit(t):
            return "mpsz"[min(t // 36, 3)]
        out = ""
        tiles = sorted(self.tiles_136)
        for su in "mpsz":
            ds = "".join(digit(t) for t in tiles if suit(t) == su)
            out += ds + su if ds else ""
        for m in self.melds:
            t0 = m.tiles[0]
            if int(m.meld_type) == 0:
                ds = "".join(digit(t) for t in m.tiles)
            else:
                ds = "0" if any(t in (16, 52, 88) for t in m.tiles) else digit(t0)
            out += f"({['', 'p', 'k', 'k', 's'][int(m.meld_type)]}{ds}{suit(t0)}0)"
        return out

    def _conditions(self, conditions):
        c = conditions or Conditions()
        if self._sanma:
            import dataclasses

            c = dataclasses.replace(c, is_sanma=True, num_players=3)
        return dataclasses_wind(c)

    def calc(self, win_tile, dora_indicators=None, conditions=None, ura_indicators=None) -> WinResult:  # hand.py:251-278
        if isinstance(conditions, (list, tuple)) and not isinstance(ura_indicators, (list, tuple)):
            # the PyO3 class orders (win_tile, dora, ura, conditions) (hand_evaluator.rs:77); accept that call shape too
            conditions, ura_indicators = ura_indicators, conditions
        tiles = self.tiles_136
        if (len(tiles) + sum(len(m.tiles) for m in self.melds)) % 3 == 1:
            tiles = sorted(tiles + [win_tile])
        q = make_query(tiles, self.melds, win_tile, dora_indicators or (), ura_indicators or (), self._conditions(conditions))
        return _to_result(eval_queries([q])[0])

    def _waits_mask(self):
        q = make_query(self.tiles_136, self.melds, 0, (), (), self._conditions(None))
        return eval_queries([q])[0].wait_mask if len(self.tiles_136) + 3 * len(self.melds) == 13 else 0

    def get_waits(self):
        m = self._waits_mask()
        return [k for k in range(34) if (m >> k) & 1]

    get_waits_u8 = get_waits

    def is_tenpai(self):
        return self._waits_mask() != 0


class HandEvaluator3P(HandEvaluator):  # hand.py:317-360 over hand_evaluator_3p.rs
    _sanma = True


def dataclasses_wind(c: Conditions) -> Conditions:
    """winds may be given as ints or Wind members (hand.py:287-296: taken mod 4)"""
    import dataclasses

    return dataclasses.replace(c, player_wind=int(c.player_wind) % 4, round_wind=int(c.round_wind) % 4)


class Score:  # score.rs:4-11
    __slots__ = ("pay_ron", "pay_tsumo_oya", "pay_tsumo_ko", "total")

    def __init__(self, pay_ron=0, pay_tsumo_oya=0, pay_tsumo_ko=0, total=0):
        self.pay_ron, self.pay_tsumo_oya, self.pay_tsumo_ko, self.total = pay_ron, pay_tsumo_oya, pay_tsumo_ko, total

    def __repr__(self):
        return (f"Score(pay_ron={self.pay_ron}, pay_tsumo_oya={self.pay_tsumo_oya}, pay_tsumo_ko={self.pay_tsumo_ko}, "
                f"total={self.total})")


def calculate_score(han, fu, is_oya, is_tsumo, honba=0, num_players=4) -> Score:  # lib.rs:5-16
    out = (C.c_uint32 * 4)()
    check(lib().rv_calculate_score(int(han), int(fu), int(bool(is_oya)), int(bool(is_tsumo)), int(honba), int(num_players), out))
    return Score(out[0], out[1], out[2], out[3])


def check_riichi_candidates(tiles_136):  # hand_evaluator.rs:263-284
    """the tiles of a 3n+2 hand whose discard leaves a tenpai hand (agari::is_tenpai of the other tiles), in input order.
    One wait query per candidate, evaluated in a single rv_hand_eval_batch; concealed hands shorter than 14 (calls made) are
    padded with far-away dummy sets, which complete nothing."""
    tiles = [int(t) for t in tiles_136]
    n = len(tiles)
    if n == 0:
        return []
    if n > 14 or n % 3 != 2:
        raise ValueError("check_riichi_candidates takes the 3n+2 concealed tiles of a hand (at most 14)")
    if any(not 0 <= t < 136 for t in tiles):
        raise ValueError("tile ids must be in 0..135")
    from .env import Meld, MeldType

    pads = [Meld(MeldType.Pon, [4 * k, 4 * k + 1, 4 * k + 2], True) for k in (27, 28, 29, 30)][: (14 - n) // 3]
    qs = [make_query(tiles[:i] + tiles[i + 1:], pads, 0, (), (), Conditions()) for i in range(n)]
    return [t for t, r in zip(tiles, eval_queries(qs)) if r.wait_mask != 0]


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
