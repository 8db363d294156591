"""Tile-string helpers of the reference's Python surface (host-side conveniences, no hot loop).

  tid <-> "1m".."9s","0m/0p/0s" (red five),"1z".."7z"   src/riichienv/convert.py:1-122
  tid <-> MJAI "5mr", "E".."C"                           riichienv-core/src/parser.rs:301-385
  parse_hand / parse_tile (text notation with melds)     riichienv-core/src/parser.rs:9-300
Tile ids: tid 0..135, kind = tid // 4; the red fives are exactly tids 16 / 52 / 88.
"""
_SUITS = "mps"
_HONORS = "ESWNPFC"
_RED = (16, 52, 88)


def tid_to_mpsz(tid: int) -> str:
    if tid in _RED:
        return f"0{_SUITS[tid // 36]}"
    if tid < 108:
        return f"{(tid % 36) // 4 + 1}{_SUITS[tid // 36]}"
    return f"{(tid - 108) // 4 + 1}z"


def tid_to_mjai(tid: int) -> str:
    if tid in _RED:
        return f"5{_SUITS[tid // 36]}r"
    if tid < 108:
        return f"{(tid % 36) // 4 + 1}{_SUITS[tid // 36]}"
    return _HONORS[(tid - 108) // 4]


def mpsz_to_tid(s: str) -> int:
    """canonical tid: copy 0 of the kind, the red copy for "0x", copy 1 for a plain five"""
    if not s:
        raise ValueError("Empty string")
    suit, num = s[-1], s[:-1]
    if suit not in "mpsz":
        raise ValueError(f"Invalid suit: {suit}")
    try:
        n = int(num)
    except ValueError as e:
        raise ValueError(f"Invalid number: {num}") from e
    if suit == "z":
        if not 1 <= n <= 7:
            raise ValueError(f"Invalid honor number: {n}")
        return 108 + 4 * (n - 1)
    base = 36 * _SUITS.index(suit)
    if n == 0:
        return base + 16
    if not 1 <= n <= 9:
        raise ValueError(f"Invalid number: {n}")
    return base + 17 if n == 5 else base + 4 * (n - 1)


def mjai_to_tid(s: str) -> int:
    if s in _HONORS and len(s) == 1:
        return 108 + 4 * _HONORS.index(s)
    if len(s) == 3 and s[2] == "r" and s[0] == "5" and s[1] in _SUITS:
        return 36 * _SUITS.index(s[1]) + 16
    if len(s) == 2 and s[0].isdigit() and s[1] in _SUITS and s[0] != "0":
        return mpsz_to_tid(s)
    raise ValueError(f"Invalid MJAI tile: {s}")


def mpsz_to_mjai(s: str) -> str:
    return tid_to_mjai(mpsz_to_tid(s))


def mjai_to_mpsz(s: str) -> str:
    return tid_to_mpsz(mjai_to_tid(s))


def tid_to_mpsz_list(tids):
    return [tid_to_mpsz(t) for t in tids]


def tid_to_mjai_list(tids):
    return [tid_to_mjai(t) for t in tids]


def _unique(bases):
    """k-th occurrence of a canonical tid -> tid + k (a red five is its own counter)"""
    seen, out = {}, []
    for b in bases:
        k = seen.get(b, 0)
        seen[b] = k + 1
        out.append(b + k)
    return out


def mpsz_to_tid_list(items):
    return _unique(mpsz_to_tid(s) for s in items)


def mjai_to_tid_list(items):
    return _unique(mjai_to_tid(s) for s in items)


def mpsz_to_mjai_list(items):
    return [mpsz_to_mjai(s) for s in items]


def mjai_to_mpsz_list(items):
    return [mjai_to_mpsz(s) for s in items]


def paishan_to_wall(paishan: str):
    """concatenated two-character tiles ("1m2m0p...") -> unique tids in the order written"""
    if len(paishan) % 2:
        raise ValueError(f"Invalid paishan string length: {len(paishan)}")
    return _unique(mpsz_to_tid(paishan[i:i + 2]) for i in range(0, len(paishan), 2))


# ---- text notation with melds (parser.rs:9-300) -------------------------------------------------
class _Copies:
    """hands out unused copies of a kind: the red copy only when asked for, plain fives from copies 1, 2, 3 first"""

    def __init__(self):
        self.used = set()

    def take(self, kind, red):
        if kind >= 34:
            raise ValueError(f"Invalid tile ID: {kind}")
        five = kind in (4, 13, 22)
        order = (0,) if (five and red) else ((1, 2, 3, 0) if five else (0, 1, 2, 3))
        for k in order:
            if (kind, k) not in self.used:
                self.used.add((kind, k))
                return 4 * kind + k
        raise ValueError(f"No more copies of tile {kind}")


def _kind_of(digit, suit):
    off = {"m": 0, "p": 9, "s": 18, "z": 27}[suit]
    d = int(digit)
    return (off + 4, True) if d == 0 else (off + d - 1, False)


def _parse_meld(body, copies):
    from .env import Meld, MeldType

    prefix = body[0] if body[:1] in ("p", "k", "s") else ""
    rest = body[len(prefix):]
    i = 0
    while i < len(rest) and rest[i].isdigit():
        i += 1
    digits, suit = rest[:i], rest[i:i + 1]
    call = int(rest[i + 1]) if len(rest) > i + 1 and rest[i + 1].isdigit() else 0
    if suit not in ("m", "p", "s", "z"):
        raise ValueError(f"Invalid suit in meld: {suit or ' '}")
    if not prefix:
        if len(digits) != 3:
            raise ValueError("Chi meld requires 3 digits")
        tiles = sorted(copies.take(*_kind_of(d, suit)) for d in digits)
        return Meld(MeldType.Chi, tiles, True, -1, None)
    kind, red = _kind_of(digits[0], suit)
    want = 3 if prefix == "p" else 4
    tiles, got_red = [], False
    if red:
        tiles.append(copies.take(kind, True))
        got_red = True
    while len(tiles) < want:
        try:
            tiles.append(copies.take(kind, False))
        except ValueError:
            if got_red:
                raise ValueError(f"Not enough tiles for meld of {kind}") from None
            try:
                tiles.append(copies.take(kind, True))
            except ValueError:
                raise ValueError(f"Not enough tiles for meld of {kind}") from None
            got_red = True
    mt = {"p": MeldType.Pon, "s": MeldType.Kakan}.get(prefix) or (MeldType.Ankan if call == 0 else MeldType.Daiminkan)
    return Meld(mt, sorted(tiles), mt != MeldType.Ankan, -1, None)


def parse_hand(text: str):
    """"123m406p(p5z1)(k2z)" -> (tids of the concealed part in the order written, melds)"""
    copies, tiles, melds, pending = _Copies(), [], [], []
    i = 0
    while i < len(text):
        c = text[i]
        if c == "(":
            j = text.find(")", i)
            j = len(text) if j < 0 else j
            melds.append(_parse_meld(text[i + 1:j], copies))
            i = j + 1
            continue
        if c.isdigit() and c.isascii():
            pending.append(c)
        elif c in "mpsz":
            for d in pending:
                tiles.append(copies.take(*_kind_of(d, c)))
            pending = []
        i += 1
    if pending:
        raise ValueError(f"Parse error in {text!r}: Pending digits without suit")
    return tiles, melds


def parse_tile(text: str) -> int:
    tiles, melds = parse_hand(text)
    if melds:
        raise ValueError("parse_tile expects a single tile, but found meld syntax in input")
    if not tiles:
        raise ValueError("No tile found in string")
    if len(tiles) != 1:
        raise ValueError(f"Expected exactly one tile, but found {len(tiles)} tiles in string")
    return tiles[0]
