"""Drop-in mirror of the reference's Python surface for the hot path, over the CUDA library.

Same names, argument meaning and error behaviour as `riichienv._riichienv` (PyO3):
  RiichiEnv        riichienv-python/src/env.rs:74-118, 353-401, 673-872
  Observation      riichienv-core/src/observation/mod.rs:24-160, observation/mjai_select.rs:87-193
  Action/ActionType/Phase   riichienv-core/src/action.rs:30-227, src/riichienv/action.py:6-17
  Meld/MeldType    riichienv-core/src/types.rs:57-108
  GameRule         riichienv-core/src/rule.rs:10-112
A single env is a VecRiichiEnv of size 1; all game logic runs in libriichienv_b200.so (no CPU fallback).
"""
import ctypes as C
import enum
import json

from . import _abi as A
from ._lib import events_to_json
from .vec_env import GAME_MODES, VecRiichiEnv


class ActionType(enum.IntEnum):  # action.rs:55-68
    DISCARD = 0
    CHI = 1
    PON = 2
    DAIMINKAN = 3
    RON = 4
    RIICHI = 5
    TSUMO = 6
    PASS = 7
    ANKAN = 8
    KAKAN = 9
    KYUSHU_KYUHAI = 10
    KITA = 11


# PascalCase aliases (src/riichienv/action.py:6-17)
for _n, _alias in [("DISCARD", "Discard"), ("CHI", "Chi"), ("PON", "Pon"), ("DAIMINKAN", "Daiminkan"), ("RON", "Ron"),
                   ("RIICHI", "Riichi"), ("TSUMO", "Tsumo"), ("PASS", "Pass"), ("ANKAN", "Ankan"), ("KAKAN", "Kakan"),
                   ("KYUSHU_KYUHAI", "KyushuKyuhai"), ("KITA", "Kita")]:
    setattr(ActionType, _alias, ActionType[_n])


class Phase(enum.IntEnum):  # action.rs:30-33
    WaitAct = 0
    WaitResponse = 1


class MeldType(enum.IntEnum):  # types.rs:57-63
    Chi = 0
    Pon = 1
    Daiminkan = 2
    Ankan = 3
    Kakan = 4


class Wind(enum.IntEnum):
    East = 0
    South = 1
    West = 2
    North = 3


class GameType(enum.IntEnum):  # src/riichienv/game_mode.py:4-10
    YON_IKKYOKU = 0
    YON_TONPUSEN = 1
    YON_HANCHAN = 2
    SAN_IKKYOKU = 3
    SAN_TONPUSEN = 4
    SAN_HANCHAN = 5


def tid_to_mjai(tid: int) -> str:  # parser.rs:301-334
    if tid == 16:
        return "5mr"
    if tid == 52:
        return "5pr"
    if tid == 88:
        return "5sr"
    if tid < 108:
        return f"{(tid % 36) // 4 + 1}{'mps'[tid // 36]}"
    return "ESWNPFC"[(tid - 108) // 4]


class Action:
    """action.rs:82-105 — consume_tiles are stored sorted."""

    __slots__ = ("action_type", "tile", "consume_tiles", "actor")

    def __init__(self, type=ActionType.PASS, tile=None, consume_tiles=(), actor=None):
        self.action_type = ActionType(int(type))
        self.tile = None if tile is None else int(tile)
        self.consume_tiles = sorted(int(t) for t in consume_tiles)
        self.actor = None if actor is None else int(actor)

    def __eq__(self, o):
        return (isinstance(o, Action) and self.action_type == o.action_type and self.tile == o.tile
                and self.consume_tiles == o.consume_tiles and self.actor == o.actor)

    def __repr__(self):
        return (f"Action(action_type={self.action_type.name}, tile={self.tile}, consume_tiles={self.consume_tiles}, "
                f"actor={self.actor})")

    def to_dict(self):
        return {"type": int(self.action_type), "tile": self.tile, "consume_tiles": list(self.consume_tiles), "actor": self.actor}

    def to_mjai(self) -> str:  # action.rs:107-150 (serde_json: keys alphabetical)
        t = {ActionType.DISCARD: "dahai", ActionType.CHI: "chi", ActionType.PON: "pon", ActionType.DAIMINKAN: "daiminkan",
             ActionType.ANKAN: "ankan", ActionType.KAKAN: "kakan", ActionType.RIICHI: "reach", ActionType.TSUMO: "hora",
             ActionType.RON: "hora", ActionType.KYUSHU_KYUHAI: "ryukyoku", ActionType.KITA: "kita",
             ActionType.PASS: "none"}[self.action_type]
        d = {"type": t}
        if self.actor is not None:
            d["actor"] = self.actor
        if self.tile is not None and self.action_type not in (ActionType.TSUMO, ActionType.RON, ActionType.RIICHI):
            d["pai"] = tid_to_mjai(self.tile)
        if self.consume_tiles:
            d["consumed"] = [tid_to_mjai(t) for t in self.consume_tiles]
        return json.dumps(d, sort_keys=True, separators=(",", ":"))

    def encode(self) -> int:  # action.rs:158-227 (82-id space)
        at = self.action_type
        if at == ActionType.DISCARD:
            if self.tile is None:
                raise ValueError("Discard action requires a tile")
            return self.tile // 4
        if at == ActionType.RIICHI:
            return 37
        if at == ActionType.CHI:
            if self.tile is None:
                raise ValueError("Chi action requires a target tile")
            kinds = sorted(set([t // 4 for t in self.consume_tiles] + [self.tile // 4]))
            if len(kinds) != 3:
                raise ValueError(f"Invalid Chi tiles: target={self.tile}, consumed={self.consume_tiles}")
            return 38 + kinds.index(self.tile // 4)
        if at == ActionType.PON:
            return 41
        if at == ActionType.DAIMINKAN:
            if self.tile is None:
                raise ValueError("Daiminkan action requires a tile")
            return 42 + self.tile // 4
        if at in (ActionType.ANKAN, ActionType.KAKAN):
            if not self.consume_tiles:
                raise ValueError("Ankan/Kakan action requires consumed tiles")
            return 42 + self.consume_tiles[0] // 4
        if at in (ActionType.RON, ActionType.TSUMO):
            return 79
        if at == ActionType.KYUSHU_KYUHAI:
            return 80
        if at == ActionType.PASS:
            return 81
        raise ValueError("Kita action is not valid in 4-player mode")

    # -- ABI marshalling
    def _to_abi(self) -> A.Action:
        a = A.Action()
        a.type = int(self.action_type)
        a.tile = 255 if self.tile is None else self.tile
        a.n_consume = min(4, len(self.consume_tiles))
        for k in range(4):
            a.consume[k] = self.consume_tiles[k] if k < a.n_consume else 255
        a.actor = 255 if self.actor is None else self.actor
        return a

    @classmethod
    def _from_abi(cls, a: A.Action) -> "Action":
        return cls(a.type, None if a.tile == 255 else a.tile, [a.consume[k] for k in range(a.n_consume)],
                   None if a.actor == 255 else a.actor)


class Action3P(Action):
    """action.rs:434-467 — the sanma wrapper: same fields, 60-id action space (ActionEncoder::encode_3p, 262-346)."""

    __slots__ = ()

    def __repr__(self):
        return (f"Action3P(action_type={self.action_type.name}, tile={self.tile}, consume_tiles={self.consume_tiles}, "
                f"actor={self.actor})")

    @staticmethod
    def _compact(tile):
        kind = tile // 4
        if kind >= 34:
            raise ValueError(f"Invalid tile type {kind} for 3P encode")
        if 1 <= kind <= 7:
            raise ValueError(f"Tile type {kind} (manzu 2-8) is not valid in 3P mode")
        return 0 if kind == 0 else kind - 7

    def encode(self) -> int:
        at = self.action_type
        if at == ActionType.DISCARD:
            if self.tile is None:
                raise ValueError("Discard action requires a tile")
            return self._compact(self.tile)
        if at == ActionType.RIICHI:
            return 27
        if at == ActionType.CHI:
            raise ValueError("Chi is not allowed in 3P mode")
        if at == ActionType.PON:
            return 28
        if at == ActionType.DAIMINKAN:
            if self.tile is None:
                raise ValueError("Daiminkan action requires a tile")
            return 29 + self._compact(self.tile)
        if at in (ActionType.ANKAN, ActionType.KAKAN):
            if not self.consume_tiles:
                raise ValueError("Ankan/Kakan action requires consumed tiles")
            return 29 + self._compact(self.consume_tiles[0])
        if at in (ActionType.RON, ActionType.TSUMO):
            return 56
        if at == ActionType.KYUSHU_KYUHAI:
            return 57
        if at == ActionType.PASS:
            return 58
        return 59  # Kita


class Meld:  # types.rs:98-190
    __slots__ = ("meld_type", "tiles", "opened", "from_who", "called_tile")

    def __init__(self, meld_type, tiles, opened, from_who=-1, called_tile=None):
        self.meld_type = MeldType(int(meld_type))
        self.tiles = [int(t) for t in tiles]
        self.opened = bool(opened)
        self.from_who = int(from_who)
        self.called_tile = called_tile

    def __repr__(self):
        return f"Meld({self.meld_type.name}, {self.tiles}, opened={self.opened}, from_who={self.from_who})"

    def __eq__(self, o):
        return isinstance(o, Meld) and (self.meld_type, self.tiles, self.opened, self.from_who) == (
            o.meld_type, o.tiles, o.opened, o.from_who)


class GameRule:  # rule.rs:10-112
    _FIELDS = ["allows_ron_on_ankan_for_kokushi_musou", "is_kokushi_musou_13machi_double", "is_suuankou_tanki_double",
               "is_junsei_chuurenpoutou_double", "is_daisuushii_double", "yakuman_pao_is_liability_only", "sanchaho_is_draw",
               "kuikae_forbidden"]

    def __init__(self, allows_ron_on_ankan_for_kokushi_musou=False, is_kokushi_musou_13machi_double=False,
                 is_suuankou_tanki_double=False, is_junsei_chuurenpoutou_double=False, is_daisuushii_double=False,
                 yakuman_pao_is_liability_only=False, sanchaho_is_draw=False, kuikae_forbidden=True):
        loc = locals()
        for f in self._FIELDS:
            setattr(self, f, bool(loc[f]))

    @staticmethod
    def default_tenhou():
        return GameRule(sanchaho_is_draw=True, kuikae_forbidden=True)

    @staticmethod
    def default_mjsoul():
        return GameRule(True, True, True, True, True, True, False, True)

    def bits(self) -> int:
        return sum((1 << i) for i, f in enumerate(self._FIELDS) if getattr(self, f))

    def __eq__(self, o):
        return isinstance(o, GameRule) and self.bits() == o.bits()

    def __repr__(self):
        return "GameRule(" + ", ".join(f"{f}={str(getattr(self, f)).lower()}" for f in self._FIELDS) + ")"


def _melds_of(s: A.GameState, p: int):
    out = []
    for m in range(s.n_melds[p]):
        tiles = [t for t in s.meld_tiles[p][m] if t != 255]
        ty = s.meld_type[p][m]
        fw = s.meld_from[p][m]
        ct = s.meld_called[p][m]
        out.append(Meld(ty, tiles, ty != MeldType.Ankan, -1 if fw == 255 else fw, None if ct == 255 else ct))
    return out


class Observation:
    """observation/mod.rs:24-160 — a by-value snapshot for one seat.  The constructor takes the reference's argument
    list (Observation::new, observation/mod.rs:62-110); observations returned by RiichiEnv are additionally bound to the
    live game so that the tensor encoders can run on the device."""

    _NP = 4

    def __init__(self, player_id, hands, melds, discards, dora_indicators, scores, riichi_declared, legal_actions, events,
                 honba=0, riichi_sticks=0, round_wind=0, oya=0, kyoku_index=0, waits=(), is_tenpai=False,
                 riichi_sutehais=None, last_tedashis=None, last_discard=None, drawn_tile=None):
        n = self._NP
        self.player_id = int(player_id)
        self.hands = [list(h) for h in hands]
        self.melds = [list(m) for m in melds]
        self.discards = [list(d) for d in discards]
        self.dora_indicators = list(dora_indicators)
        self.scores = list(scores)
        self.riichi_declared = [bool(x) for x in riichi_declared]
        self._legal_actions = list(legal_actions)
        self._new_events = list(events)
        self.honba, self.riichi_sticks, self.round_wind, self.oya = honba, riichi_sticks, round_wind, oya
        self.kyoku_index = kyoku_index
        self.waits = list(waits)
        self.is_tenpai = bool(is_tenpai)
        self.tsumogiri_flags = [[] for _ in range(n)]  # always empty in the live env (observation/mod.rs:105)
        self.riichi_sutehais = list(riichi_sutehais) if riichi_sutehais is not None else [None] * n
        self.last_tedashis = list(last_tedashis) if last_tedashis is not None else [None] * n
        self.last_discard = last_discard
        self.drawn_tile = drawn_tile
        self._env = None
        self._token = None
        self._seq_start_word = 0
        self._seq = None

    @classmethod
    def _from_state(cls, env: "RiichiEnv", s: A.GameState, pid: int, legal, new_events, seq_start_word=0):
        n = env._np
        waits = [k for k in range(34) if (s.c_waits[pid] >> k) & 1]
        o = cls(pid,
                [[s.hand[p][k] for k in range(s.hand_len[p])] if p == pid else [] for p in range(n)],
                [_melds_of(s, p) for p in range(n)],
                [[s.river[p][k] for k in range(min(s.n_river[p], A.RIVER_CAP))] for p in range(n)],
                [s.dora_ind[k] for k in range(s.n_dora)], [s.score[p] for p in range(n)],
                [bool(s.flags[p] & A.F_RIICHI_DECLARED) for p in range(n)], legal, new_events,
                s.honba, s.riichi_sticks, s.round_wind, s.oya, s.kyoku_idx, waits, bool(waits),
                [None if s.riichi_sutehai[p] == 255 else s.riichi_sutehai[p] for p in range(n)],
                [None if s.last_tedashi[p] == 255 else s.last_tedashi[p] for p in range(n)],
                # state/mod.rs:252 destructures (pid, tile) as (tile, _pid): the field carries the DISCARDER'S SEAT
                None if s.last_discard_pid == 255 else s.last_discard_pid,
                None if s.drawn_tile == 255 else s.drawn_tile)
        o._env, o._token, o._seq_start_word = env, env._token, seq_start_word
        return o

    def _live(self, what):
        if self._env is None or self._token != self._env._token:
            raise RuntimeError(f"{what} is computed on the device from the live game: call it on the observations of the "
                               "latest reset()/step()")
        return self._env

    # ---- serde (observation/mod.rs:143-160): base64 of the struct's JSON, field names as serde derives them ----
    def serialize_to_base64(self) -> str:
        import base64

        def meld(m):
            return {"meld_type": m.meld_type.name, "tiles": m.tiles, "opened": m.opened, "from_who": m.from_who,
                    "called_tile": m.called_tile}

        def act(a):
            return {"action_type": _ACTION_SERDE[a.action_type], "tile": a.tile, "consume_tiles": a.consume_tiles, "actor": a.actor}
        d = {"player_id": self.player_id, "hands": self.hands, "melds": [[meld(m) for m in ms] for ms in self.melds],
             "discards": self.discards, "dora_indicators": self.dora_indicators, "scores": self.scores,
             "riichi_declared": self.riichi_declared, "_legal_actions": [act(a) for a in self._legal_actions],
             "events": self._new_events, "honba": self.honba, "riichi_sticks": self.riichi_sticks, "round_wind": self.round_wind,
             "oya": self.oya, "kyoku_index": self.kyoku_index, "waits": self.waits, "is_tenpai": self.is_tenpai,
             "tsumogiri_flags": self.tsumogiri_flags, "riichi_sutehais": self.riichi_sutehais,
             "last_tedashis": self.last_tedashis, "last_discard": self.last_discard, "drawn_tile": self.drawn_tile}
        return base64.b64encode(json.dumps(d, separators=(",", ":")).encode()).decode()

    @classmethod
    def deserialize_from_base64(cls, text: str):
        import base64
        import binascii

        try:
            raw = base64.b64decode(text, validate=True)
        except (binascii.Error, ValueError) as e:
            raise ValueError(f"Serialization error: base64 decode failed: {e}") from None
        try:
            d = json.loads(raw)
            names = {v: k for k, v in _ACTION_SERDE.items()}
            acts = [cls._ACTION(names[a["action_type"]], a["tile"], a["consume_tiles"], a["actor"]) for a in d["_legal_actions"]]
            melds = [[Meld(MeldType[m["meld_type"]], m["tiles"], m["opened"], m["from_who"], m.get("called_tile")) for m in ms]
                     for ms in d["melds"]]
            o = cls(d["player_id"], d["hands"], melds, d["discards"], d["dora_indicators"], d["scores"], d["riichi_declared"],
                    acts, d["events"], d["honba"], d["riichi_sticks"], d["round_wind"], d["oya"], d["kyoku_index"], d["waits"],
                    d["is_tenpai"], d["riichi_sutehais"], d["last_tedashis"], d["last_discard"], d.get("drawn_tile"))
            o.tsumogiri_flags = d["tsumogiri_flags"]
            return o
        except (KeyError, TypeError, ValueError) as e:
            raise ValueError(f"Serialization error: JSON deserialize failed: {e}") from None

    @property
    def hand(self):
        return list(self.hands[self.player_id])

    @property
    def events(self):
        """the event delta of this observation as parsed objects (observation/python.rs:81-92)"""
        return [json.loads(x) for x in self._new_events]

    def legal_actions(self):
        return list(self._legal_actions)

    def new_events(self):
        return list(self._new_events)

    @property
    def action_space_size(self):  # observation/python.rs:113-116, observation_3p/python.rs:116-118
        return 60 if self._NP == 3 else 82

    def mask(self):  # observation/python.rs:98-111, observation_3p/python.rs:102-114
        m = bytearray(self.action_space_size)
        for a in self._legal_actions:
            try:
                m[a.encode()] = 1
            except ValueError:
                pass
        return bytes(m)

    def find_action(self, action_id: int):  # observation/mod.rs:117-129
        for a in self._legal_actions:
            try:
                if a.encode() == action_id:
                    return a
            except ValueError:
                pass
        return None

    def select_action_from_mjai(self, mjai):  # observation/mjai_select.rs:19-193
        if isinstance(mjai, str):
            try:
                v = json.loads(mjai)
            except ValueError:
                return None
            tile_str = v.get("pai") or ""
        elif isinstance(mjai, dict):
            v = mjai
            tile_str = v.get("pai") or v.get("tile") or ""
        else:
            return None
        atype = v.get("type", "")
        tsumogiri = v.get("tsumogiri") if isinstance(v.get("tsumogiri"), bool) else None
        consumed = v.get("consumed") if isinstance(v.get("consumed"), list) else None
        la = self._legal_actions
        if atype == "hora":
            return next((a for a in la if a.action_type in (ActionType.TSUMO, ActionType.RON)), None)
        if atype == "none":
            return next((a for a in la if a.action_type == ActionType.PASS), None)
        tt = {"dahai": ActionType.DISCARD, "chi": ActionType.CHI, "pon": ActionType.PON, "kakan": ActionType.KAKAN,
              "daiminkan": ActionType.DAIMINKAN, "ankan": ActionType.ANKAN, "reach": ActionType.RIICHI,
              "ryukyoku": ActionType.KYUSHU_KYUHAI}.get(atype)
        if self._NP == 3:                       # mjai_select.rs:112-119: sanma knows `kita` and has no chi
            tt = ActionType.KITA if atype == "kita" else None if atype == "chi" else tt
        if tt is None:
            return None
        if tt == ActionType.DISCARD:
            cands = [a for a in la if a.action_type == ActionType.DISCARD and
                     (not tile_str or (a.tile is not None and tid_to_mjai(a.tile) == tile_str))]
            if not cands:
                return None
            if tsumogiri is not None and self.drawn_tile is not None:
                for a in cands:
                    if (a.tile == self.drawn_tile) == tsumogiri:
                        return a
            return cands[0]
        for a in la:
            if a.action_type != tt:
                continue
            if consumed is not None:
                if sorted(tid_to_mjai(t) for t in a.consume_tiles) != sorted(consumed):
                    continue
                if tile_str and tt in (ActionType.CHI, ActionType.PON, ActionType.DAIMINKAN, ActionType.KAKAN):
                    if a.tile is None or tid_to_mjai(a.tile) != tile_str:
                        continue
                return a
            if tile_str:
                if a.tile is not None and tid_to_mjai(a.tile) == tile_str:
                    return a
                continue
            return a
        return None

    def to_dict(self):
        return {"player_id": self.player_id, "hands": self.hands, "discards": self.discards,
                "dora_indicators": self.dora_indicators, "scores": self.scores, "riichi_declared": self.riichi_declared,
                "legal_actions": [a.to_dict() for a in self._legal_actions], "events": self.new_events(), "honba": self.honba,
                "riichi_sticks": self.riichi_sticks, "round_wind": self.round_wind, "oya": self.oya}

    def encode(self):
        """(74, 34) float32 FEATURE_ENCODING tensor bytes (observation/python.rs:457-806); sanma: (74, 27)
        (observation_3p/python.rs:402-708).  Computed on the GPU."""
        return self._live("encode()")._encode(self.player_id)

    def encode_extended(self):  # observation/python.rs:1272-1294 (4P)
        return self._ext_row().tobytes()

    def _ext_row(self):
        """(215, 34) float32 array of this observation, computed once by obs_ext_kernel (rv_vec_encode_ext)."""
        import numpy as np

        if getattr(self, "_ext", None) is None:
            env = self._live("encode_extended()")
            self._ext = np.frombuffer(env._encode(self.player_id, extended=True), dtype=np.float32).reshape(215, -1)
        return self._ext

    # The standalone encoders of observation/python.rs:195-1270 (sanma: observation_3p/python.rs:198-1115) are the channel
    # blocks of encode_extended — encode.rs holds the same bodies as `_into` variants; a statement-level diff shows only
    # renames — so they are slices of the device-computed row.  Sanma rows carry three relative seats per group and two
    # opponents (NP = 3): the slices stop there.
    def _seats(self):
        return self._NP

    def encode_discard_history_decay(self, decay_rate=None):  # python.rs:196-249 -> (NP, W)
        if decay_rate is not None and float(decay_rate) != 0.2:
            raise NotImplementedError("the device table holds exp(-0.2 * age): only the default decay_rate")
        return self._ext_row()[74:74 + self._seats()].tobytes()

    def encode_shanten_efficiency(self):  # python.rs:820-878 -> (NP, 4)
        return self._ext_row()[78:78 + 4 * self._seats(), 0].tobytes()

    def encode_ankan_overview(self):  # python.rs:976-1010 -> (NP, W)
        return self._ext_row()[94:94 + self._seats()].tobytes()

    def encode_fuuro_overview(self):  # python.rs:931-974 -> (NP, 4, 5, W)
        return self._ext_row()[98:98 + 20 * self._seats()].tobytes()

    def encode_action_availability(self):  # python.rs:1012-1065 -> (11,)
        return self._ext_row()[178:189, 0].tobytes()

    def encode_discard_candidates(self):  # python.rs:1208-1270 -> (5,)
        return self._ext_row()[189:194, 0].tobytes()

    def encode_pass_context(self):  # python.rs:1165-1206 -> (3,)
        return self._ext_row()[194:197, 0].tobytes()

    def encode_last_tedashis(self):  # python.rs:1117-1163 -> (NP - 1, 3)
        return self._ext_row()[197:197 + 3 * (self._seats() - 1), 0].tobytes()

    def encode_riichi_sutehais(self):  # python.rs:1067-1115 -> (NP - 1, 3)
        return self._ext_row()[206:206 + 3 * (self._seats() - 1), 0].tobytes()

    def encode_yaku_possibility(self):  # observation/python.rs:327-449, observation_3p/python.rs:275-395 -> (NP, 21, 2)
        from . import yaku_possibility

        return yaku_possibility.encode(self, self._NP)

    def encode_furiten_ron_possibility(self):  # python.rs:251-293 -> (NP, 21)
        """All ones: the encoder only clears a seat's row after three consecutive tsumogiri flags, and the live env never
        fills `tsumogiri_flags` (observation/mod.rs:105) — a constant, so nothing is computed."""
        import numpy as np

        return np.ones((self._seats(), 21), np.float32).tobytes()

    def encode_kawa_overview(self):  # python.rs:881-930 -> (4, 7, 34), seats in absolute order (obs_kawa_kernel; 4P)
        return self._live("encode_kawa_overview()")._v.encode_kawa_single(self.player_id)

    # ---- sequence features (observation/python.rs:1297-1364): raw bytes, as the reference returns them ----
    def _seq_features(self):
        if self._seq is None:
            self._seq = self._live("sequence features")._encode_seq(self.player_id, self._seq_start_word)
        return self._seq

    def encode_seq_sparse(self, game_style=1):
        sp, _, _, _, lens = self._seq_features()
        sp = sp.copy()
        sp[0] = min(int(game_style), 1)
        return sp[: lens[0]].tobytes()

    def encode_seq_numeric(self):
        return self._seq_features()[1].tobytes()

    def encode_seq_progression(self):
        _, _, pr, _, lens = self._seq_features()
        return pr[: lens[1]].tobytes()

    def encode_seq_candidates(self):
        _, _, _, ca, lens = self._seq_features()
        return ca[: lens[2]].tobytes()


class Observation3P(Observation):
    """observation_3p/mod.rs:19-46: the same snapshot over 3 seats, 27 tile columns and the 60-id action space"""

    _NP = 3


_ACTION_SERDE = {ActionType.DISCARD: "Discard", ActionType.CHI: "Chi", ActionType.PON: "Pon", ActionType.DAIMINKAN: "Daiminkan",
                 ActionType.RON: "Ron", ActionType.RIICHI: "Riichi", ActionType.TSUMO: "Tsumo", ActionType.PASS: "Pass",
                 ActionType.ANKAN: "Ankan", ActionType.KAKAN: "Kakan", ActionType.KYUSHU_KYUHAI: "KyushuKyuhai",
                 ActionType.KITA: "Kita"}
Observation._ACTION = Action
Observation3P._ACTION = Action3P


class RiichiEnv:
    """env.rs:74-118 / 799-872 over a VecRiichiEnv of one game."""

    def __init__(self, game_mode=None, skip_mjai_logging=False, seed=None, round_wind=None, rule=None, device=0):
        if game_mode is None:
            gm = 0
        elif isinstance(game_mode, str):
            gm = GAME_MODES.get(game_mode, 0)  # unknown string silently -> 0 (env.rs:100)
        else:
            gm = int(game_mode)
        self.game_mode = gm
        self._np = 3 if gm >= 3 else 4
        self.skip_mjai_logging = bool(skip_mjai_logging)
        self.rule = rule or GameRule.default_tenhou()
        if seed is None:
            import os

            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = int(seed)
        self._round_wind0 = 0 if round_wind is None else int(round_wind)
        self._v = VecRiichiEnv(1, gm, self.rule.bits(), seeds=[self.seed], log_cap_words=0 if skip_mjai_logging else 1 << 16,
                               device=device)
        self._event_counts = [0, 0, 0, 0]
        self._log_cache = {}
        self._ext_log = None
        self._token = getattr(self, "_token", 0) + 1
        # GameState::new deals a first round immediately (state/mod.rs:165); mirror it so getters work before reset().
        # That constructor deal uses shuffle #0; the library's create already accounts for it (hand_index = 1), so we
        # temporarily rewind to reproduce the same wall, as a freshly constructed reference env would show.
        s = self._v.get_state(0)
        s.hand_index = 0
        self._v.set_state(0, s)
        self._v.reset(round_wind=self._round_wind0)

    # ---- core API -------------------------------------------------------------------------------
    def reset(self, oya=None, wall=None, round_wind=None, scores=None, honba=None, kyotaku=None, seed=None):
        if scores is not None and len(scores) != self._np:
            raise ValueError(f"scores length {len(scores)} does not match number of players {self._np}")
        # reset(seed=) sets GameState.seed which nothing reads (env.rs:835-837): the wall is NOT reseeded.
        self._event_counts = [0, 0, 0, 0]
        self._log_cache = {}
        self._ext_log = None
        self._token = getattr(self, "_token", 0) + 1
        self._v.reset(oya=0 if oya is None else oya, round_wind=0 if round_wind is None else round_wind,
                      honba=0 if honba is None else honba, kyotaku=0 if kyotaku is None else kyotaku,
                      scores=None if scores is None else [list(scores)], walls=None if wall is None else [self._fit_wall(wall)])
        return self._observations(self.active_players)

    def _fit_wall(self, wall):
        """reset(wall=) -> load_wall (state/wall.rs:69-80) takes any Vec; the record holds the 136 (108) tiles of a real wall.
        A longer list keeps its first tiles — the ones dealt and drawn first (the reference's tests pass 200-entry walls and
        only ever reach the first few dozen); a shorter one is refused."""
        need = 108 if self._np == 3 else 136
        wall = [int(t) for t in wall]
        if len(wall) < need:
            raise ValueError(f"wall holds {len(wall)} tiles, the {self._np}-player game needs {need}")
        return wall[:need]

    def step(self, actions):
        self._token += 1
        arr = (A.Action * 4)()
        for p in range(4):
            arr[p].type = A.NO_ACTION
        for p in range(self._np):
            a = actions.get(p) if actions else None
            if a is None:
                arr[p].type = A.NO_ACTION
            else:
                arr[p] = a._to_abi()
        self._v.step(arr)
        s = self._state()
        if s.last_error != 255:
            return {}
        return self._observations(self._active(s), s)

    def done(self):
        return bool(self._state().is_done)

    def scores(self):
        s = self._state()
        return [s.score[p] for p in range(self._np)]

    def ranks(self):  # env.rs:673-689
        sc = self.scores()
        order = sorted(range(self._np), key=lambda p: (-sc[p], p))
        out = [0] * self._np
        for r, p in enumerate(order):
            out[p] = r + 1
        return out

    def points(self, rule_name: str):  # env.rs:691-727
        if self._np == 3:
            presets = {"basic": (1.0, 35000.0, [40.0, 0.0, -40.0])}
        else:
            presets = {"basic": (1.0, 25000.0, [50.0, 10.0, -10.0, -50.0]),
                       "ouza-tyoujyo": (0.0, 25000.0, [100.0, 40.0, -40.0, -100.0]),
                       "ouza-normal": (0.0, 25000.0, [50.0, 20.0, -20.0, -50.0])}
        if rule_name not in presets:
            raise ValueError(f"Unknown preset rule{' for 3P' if self._np == 3 else ''}: {rule_name}")
        w, base, uma = presets[rule_name]
        sc, rk = self.scores(), self.ranks()
        return [(sc[i] - base) / 1000.0 * w + uma[rk[i] - 1] for i in range(self._np)]

    def get_observation(self, player_id: int):
        return self._observations([player_id])[player_id]

    def get_observations(self, players=None):
        return self._observations(list(range(self._np)) if players is None else list(players))

    get_obs_py = get_observations   # env.rs:780 (private in the reference, used by one of its tests)

    def _get_legal_actions(self, pid: int):
        acts, counts = self._v.legal_actions()
        return [self._action_cls._from_abi(acts[pid * A.MAX_LEGAL + k]) for k in range(int(counts[0, pid]))]

    # ---- MJAI-driven state tracking (env.rs:880-948 over state/event_handler.rs:18-330) -------------
    _EVENT_TYPES = {"start_game": A.EV_START_GAME, "start_kyoku": A.EV_START_KYOKU, "tsumo": A.EV_TSUMO, "dahai": A.EV_DAHAI,
                    "pon": A.EV_PON, "chi": A.EV_CHI, "kan": A.EV_DAIMINKAN, "daiminkan": A.EV_DAIMINKAN, "kakan": A.EV_KAKAN,
                    "ankan": A.EV_ANKAN, "dora": A.EV_DORA, "reach": A.EV_REACH, "reach_accepted": A.EV_REACH_ACCEPTED,
                    "hora": A.EV_HORA, "ryukyoku": A.EV_RYUKYOKU, "kita": A.EV_KITA, "end_game": A.EV_END_GAME,
                    "end_kyoku": A.EV_END_KYOKU}
    _NO_DECISION = ("start_game", "start_kyoku", "reach_accepted", "dora", "hora", "ryukyoku", "end_kyoku", "end_game")

    @staticmethod
    def _tile(s):
        """parse_mjai_tile (event_handler.rs:8-10): mjai_to_tid, an unparsable tile (\"?\") is 0"""
        from .convert import mjai_to_tid

        try:
            return mjai_to_tid(s)
        except (ValueError, TypeError):
            return 0

    def _parse_event(self, event):
        """Python dict -> (rv_mjai_event, canonical JSON text); what serde does with MjaiEvent (replay/mjai_replay.rs:67-156):
        required fields missing or of the wrong type raise ValueError, unknown types are MjaiEvent::Other (ignored)."""
        try:
            text = json.dumps(event, sort_keys=True, separators=(",", ":"), ensure_ascii=False)
        except (TypeError, ValueError) as e:
            raise ValueError(f"JSON Parse Error: {e}") from None
        ev = json.loads(text)
        if not isinstance(ev, dict) or not isinstance(ev.get("type"), str):
            raise ValueError("JSON Parse Error: missing field `type`")
        kind = ev["type"]
        e = A.MjaiEvent()
        e.type = self._EVENT_TYPES.get(kind, 0)

        def need(name, ty):
            v = ev.get(name)
            if not isinstance(v, ty) or isinstance(v, bool) and ty is int:
                raise ValueError(f"JSON Parse Error: missing field `{name}`")
            return v

        def seat(name):
            v = need(name, int)
            if v < 0:
                raise ValueError(f"JSON Parse Error: invalid value for `{name}`")
            return min(v, 255)

        def consumed():
            c = need("consumed", list)
            e.n_consumed = min(4, len(c))
            for k in range(e.n_consumed):
                e.consumed[k] = self._tile(c[k])

        if kind == "start_kyoku":
            e.bakaze = {"E": 0, "S": 1, "W": 2, "N": 3}.get(need("bakaze", str), 0)
            e.kyoku, e.honba, e.oya = need("kyoku", int) & 0xFF, need("honba", int) & 0xFF, need("oya", int) & 0xFF
            ky = ev.get("kyoutaku", ev.get("kyotaku"))
            if not isinstance(ky, int):
                raise ValueError("JSON Parse Error: missing field `kyoutaku`")
            e.kyotaku = ky & 0xFF                      # serde field type u8
            scores, tehais = need("scores", list), need("tehais", list)
            e.dora_marker = self._tile(need("dora_marker", str))
            for p in range(min(4, len(scores))):
                e.scores[p] = int(scores[p])
            for p in range(min(4, len(tehais))):
                e.tehai_len[p] = min(14, len(tehais[p]))
                for k in range(e.tehai_len[p]):
                    e.tehais[p][k] = self._tile(tehais[p][k])
            if len(scores) < self._np or len(tehais) < self._np:
                raise ValueError("start_kyoku needs scores and tehais for every seat")
        elif kind in ("tsumo", "kakan"):
            e.actor, e.pai = seat("actor"), self._tile(need("pai", str))
        elif kind == "dahai":
            e.actor, e.pai = seat("actor"), self._tile(need("pai", str))
            if need("tsumogiri", bool):
                e.type = A.EV_DAHAI_TSUMOGIRI
        elif kind in ("pon", "chi", "kan", "daiminkan"):
            e.actor, e.target, e.pai = seat("actor"), seat("target"), self._tile(need("pai", str))
            consumed()
            if kind != "kan" and kind != "daiminkan" and e.n_consumed < 2:
                raise IndexError("consumed needs two tiles")
        elif kind == "ankan":
            e.actor = seat("actor")
            consumed()
        elif kind == "dora":
            e.pai = self._tile(need("dora_marker", str))
        elif kind in ("reach", "reach_accepted", "kita"):
            e.actor = seat("actor")
        elif kind == "hora":
            e.actor, e.target = seat("actor"), seat("target")
        if kind in ("tsumo", "dahai", "pon", "chi", "kan", "daiminkan", "kakan", "ankan", "reach", "reach_accepted", "kita") \
                and e.actor >= self._np:
            raise IndexError(f"actor {e.actor} out of range")
        return e, text, kind

    def _apply(self, event):
        e, text, kind = self._parse_event(event)
        if kind == "start_game":          # env.rs:56-68: logs and counters restart with the caller's own start_game line
            self._ext_log = []
            self._event_counts = [0, 0, 0, 0]
            self._log_cache = {}
        elif self._ext_log is None:        # events fed on top of a played game: the text log continues from what was played
            played = [self._log_text(v) for v in [-1] + list(range(self._np))]
            self._ext_log = [[played[v][k] for v in range(self._np + 1)] for k in range(len(played[0]))]
        self._token += 1
        self._v.apply_events((A.MjaiEvent * 1)(e))
        self._ext_log.append(self._masked_variants(text, kind))
        return kind

    def _masked_variants(self, text, kind):
        """[all-seeing, seat 0, seat 1, ...] renderings of one fed event: _push_mjai_event's masking (state/mod.rs:2109-2143)"""
        out = [text]
        ev = json.loads(text)
        for pid in range(self._np):
            m = text
            if kind == "start_kyoku" and isinstance(ev.get("tehais"), list):
                d = dict(ev)
                d["tehais"] = [h if i == pid else ["?"] * (len(h) if isinstance(h, list) else 13) for i, h in enumerate(ev["tehais"])]
                m = json.dumps(d, sort_keys=True, separators=(",", ":"), ensure_ascii=False)
            elif kind == "tsumo" and isinstance(ev.get("actor"), int) and ev["actor"] != pid:
                d = dict(ev)
                d["pai"] = "?"
                m = json.dumps(d, sort_keys=True, separators=(",", ":"), ensure_ascii=False)
            out.append(m)
        return out

    def apply_event(self, event):
        """Apply one MJAI event (dict) to the tracked state without returning an observation (env.rs:880-887)."""
        self._apply(event)

    def observe_event(self, event, player_id):
        """Apply one MJAI event and return `player_id`'s observation if that seat now has legal actions, else None
        (env.rs:894-948)."""
        kind = self._apply(event)
        if kind in self._NO_DECISION or kind not in self._EVENT_TYPES:
            return None
        obs = self._observations([int(player_id)])[int(player_id)]
        return obs if obs.legal_actions() else None

    # ---- logs -----------------------------------------------------------------------------------
    @property
    def mjai_log(self):
        if self.skip_mjai_logging:
            return []
        return [json.loads(x) for x in self._log_text(-1)]

    @mjai_log.setter
    def mjai_log(self, v):     # tests/env/helper.py assigns a list of dicts; the reference's setter takes JSON objects
        self._ext_log = [self._masked_variants(json.dumps(e, sort_keys=True, separators=(",", ":"), ensure_ascii=False),
                                               e.get("type", "") if isinstance(e, dict) else "") for e in v]

    def _masked_log(self, pid):
        if self.skip_mjai_logging:
            return []
        return self._log_text(pid)

    def _log_text(self, viewer):
        """rendered log of one viewer (-1: the all-seeing log); only the events pushed since the last call are rendered"""
        if self._ext_log is not None:      # event-driven mode: the log is the text that was fed (apply_event keeps it)
            return [row[viewer + 1] for row in self._ext_log]
        have = self._log_cache.setdefault(viewer, [])
        have.extend(self._v.mjai_log(0, viewer, skip_events=len(have)))
        return have

    # ---- internals ------------------------------------------------------------------------------
    def _state(self) -> A.GameState:
        return self._v.get_state(0)

    @staticmethod
    def _active(s):
        return [p for p in range(4) if (s.active_mask >> p) & 1]

    def _observations(self, players, s=None):
        s = s or self._state()
        acts, counts = self._v.legal_actions()
        out = {}
        logs = {}
        for p in players:
            legal = [self._action_cls._from_abi(acts[p * A.MAX_LEGAL + k]) for k in range(int(counts[0, p]))]
            full = logs.setdefault(p, self._masked_log(p))
            new = full[self._event_counts[p]:]
            first_new = self._event_counts[p]   # (its word offset in the binary log is only looked up if sequence features are asked for)
            self._event_counts[p] = len(full)  # state/mod.rs:211-218: the delta advances on every observation
            out[p] = (Observation3P if self._np == 3 else Observation)._from_state(self, s, p, legal, new, first_new)
        return out

    def _event_word_offset(self, k):
        """word offset of the k-th event of this game's binary log (every seat's masked log has the same event count)"""
        if self.skip_mjai_logging or k == 0:
            return 0
        words = self._v.events(0)
        i = 0
        for _ in range(k):
            i += max(1, (int(words[i]) >> 8) & 0xFF)
        return i

    def _encode_seq(self, pid, first_new_event):
        if self.skip_mjai_logging:
            raise ValueError("sequence features need the MJAI log (skip_mjai_logging=False)")
        if self._ext_log is not None:
            raise NotImplementedError("sequence features read the device event log; a game fed through apply_event has none")
        return self._v.encode_seq_single(pid, self._event_word_offset(first_new_event))

    def _encode(self, pid, extended=False):
        """bytes of the (74, 34) — sanma (74, 27) — float32 tensor for seat `pid` (must owe an action), computed by
        obs_encode_kernel; extended=True: the (215, W) tensor of encode_extended (obs_ext_kernel)."""
        return self._v.encode_single(pid, extended)

    # ---- getters / setters used by callers and by the reference's tests (env.rs:134-635) --------------
    # Every property reads / writes the game's snapshot record (rv_vec_get_state / rv_vec_set_state); the library
    # recomputes the derived caches of a record it is handed.  Lists are by-value copies, as in PyO3: mutate and assign back.
    num_players = property(lambda self: self._np)
    action_space_size = property(lambda self: 60 if self._np == 3 else 82)
    _action_cls = property(lambda self: Action3P if self._np == 3 else Action)
    kyoku_idx = property(lambda self: self._state().kyoku_idx)
    _custom_round_wind = property(lambda self: self._state().round_wind)   # env.rs:633
    last_error = property(lambda self: None if self._state().last_error == 255 else
                          f"Error: Illegal Action by Player {self._state().last_error}")
    score_deltas = property(lambda self: [self._state().score_delta[p] for p in range(self._np)])   # env.rs:611

    def _mutate(self, fn):
        s = self._state()
        fn(s)
        self._v.set_state(0, s)

    def _scalar(name, conv=int, get=None):  # noqa: N805 — property factory, evaluated in the class body
        def fget(self):
            v = getattr(self._state(), name)
            return get(v) if get else v

        def fset(self, v):
            self._mutate(lambda s: setattr(s, name, conv(v)))
        return property(fget, fset)

    def _seat_flag(bit):  # noqa: N805
        def fget(self):
            s = self._state()
            return [bool(s.flags[p] & bit) for p in range(self._np)]

        def fset(self, v):
            if len(v) != self._np:      # the PyO3 setters ignore a list of the wrong length (env.rs:450-459)
                return

            def f(s):
                for p in range(self._np):
                    s.flags[p] = (s.flags[p] | bit) if v[p] else (s.flags[p] & ~bit)
            self._mutate(f)
        return property(fget, fset)

    def _opt_u8(v):  # noqa: N805
        return 255 if v is None else int(v)

    current_player = _scalar("current_player")
    oya = _scalar("oya")
    honba = _scalar("honba")
    round_wind = _scalar("round_wind")
    riichi_sticks = _scalar("riichi_sticks")
    turn_count = _scalar("turn_count")
    pending_kan_dora_count = _scalar("pending_kan_dora_count")
    is_done = _scalar("is_done", lambda v: int(bool(v)), bool)
    needs_tsumo = _scalar("needs_tsumo", lambda v: int(bool(v)), bool)
    is_first_turn = _scalar("is_first_turn", lambda v: int(bool(v)), bool)
    is_rinshan_flag = _scalar("is_rinshan_flag", lambda v: int(bool(v)), bool)
    drawn_tile = _scalar("drawn_tile", _opt_u8, lambda t: None if t == 255 else t)
    riichi_declared = _seat_flag(A.F_RIICHI_DECLARED)
    riichi_stage = _seat_flag(A.F_RIICHI_STAGE)
    missed_agari_doujun = _seat_flag(A.F_MISSED_AGARI_DOUJUN)
    # state/mod.rs:84,335-338: armed by nothing in the live 4P env and consumed at the top of step(); the records carry no
    # such flag because a finished round is re-dealt inside the step that ends it
    needs_initialize_next_round = property(lambda self: False, lambda self, v: None)

    @property
    def phase(self):
        return Phase(self._state().phase)

    @phase.setter
    def phase(self, v):  # env.rs:482-498: Phase or int, unknown ints -> WaitAct
        if not isinstance(v, (int, Phase)):
            raise TypeError("Expected Phase or int")
        self._mutate(lambda s: setattr(s, "phase", 1 if int(v) == 1 else 0))

    @property
    def active_players(self):
        return self._active(self._state())

    @active_players.setter
    def active_players(self, v):
        self._mutate(lambda s: setattr(s, "active_mask", sum(1 << int(p) for p in set(v))))

    @property
    def hands(self):
        s = self._state()
        return [[s.hand[p][k] for k in range(s.hand_len[p])] for p in range(self._np)]

    @hands.setter
    def hands(self, v):
        if len(v) != self._np:
            return

        def f(s):
            for p in range(self._np):
                tiles = [int(t) for t in v[p]]
                if len(tiles) > A.HAND_CAP:
                    raise ValueError(f"a hand holds at most {A.HAND_CAP} tiles, got {len(tiles)}")
                for k in range(A.HAND_CAP):
                    s.hand[p][k] = tiles[k] if k < len(tiles) else 255
                s.hand_len[p] = len(tiles)
        self._mutate(f)

    @property
    def melds(self):
        s = self._state()
        return [_melds_of(s, p) for p in range(self._np)]

    @melds.setter
    def melds(self, v):
        if len(v) != self._np:
            return

        def f(s):
            for p in range(self._np):
                ms = list(v[p])
                if len(ms) > 4:
                    raise ValueError(f"a seat holds at most 4 melds, got {len(ms)}")
                s.n_melds[p] = len(ms)
                for m in range(4):          # the storage is 4 melds x 4 tiles whatever the seat count
                    for k in range(4):
                        s.meld_tiles[p][m][k] = 255
                    s.meld_type[p][m] = s.meld_from[p][m] = s.meld_called[p][m] = 255
                for m, md in enumerate(ms):
                    for k, t in enumerate(md.tiles[:4]):
                        s.meld_tiles[p][m][k] = t
                    s.meld_type[p][m] = int(md.meld_type)
                    s.meld_from[p][m] = 255 if md.from_who < 0 else md.from_who
                    s.meld_called[p][m] = 255 if md.called_tile is None else md.called_tile
        self._mutate(f)

    @property
    def discards(self):
        s = self._state()
        return [[s.river[p][k] for k in range(min(s.n_river[p], A.RIVER_CAP))] for p in range(self._np)]

    @discards.setter
    def discards(self, v):
        """PlayerState.discards only (env.rs:193-202): the parallel discard_from_hand / discard_is_riichi vectors keep
        their contents — bits past the new length are dead storage, bits below it survive."""
        if len(v) != self._np:
            return

        def f(s):
            for p in range(self._np):
                d = [int(t) for t in v[p]]
                if len(d) > A.RIVER_CAP:
                    raise ValueError(f"a river holds at most {A.RIVER_CAP} discards, got {len(d)}")
                s.n_river[p] = len(d)
                for k in range(A.RIVER_CAP):
                    s.river[p][k] = d[k] if k < len(d) else 255
        self._mutate(f)

    def _river_bits(name):  # noqa: N805
        def fget(self):
            s = self._state()
            return [[bool((getattr(s, name)[p] >> k) & 1) for k in range(min(s.n_river[p], A.RIVER_CAP))] for p in range(self._np)]

        def fset(self, v):
            if len(v) != self._np:
                return

            def f(s):
                for p in range(self._np):
                    getattr(s, name)[p] = sum(1 << k for k, b in enumerate(list(v[p])[: A.RIVER_CAP]) if b)
            self._mutate(f)
        return property(fget, fset)

    discard_from_hand = _river_bits("river_tedashi")      # env.rs:205-222
    discard_is_riichi = _river_bits("river_riichi")       # env.rs:225-242

    @property
    def riichi_declaration_index(self):                   # env.rs:287-304
        s = self._state()
        return [None if s.riichi_decl_idx[p] == 255 else s.riichi_decl_idx[p] for p in range(self._np)]

    @riichi_declaration_index.setter
    def riichi_declaration_index(self, v):
        if len(v) == self._np:
            self._mutate(lambda s: [s.riichi_decl_idx.__setitem__(p, 255 if v[p] is None else int(v[p])) for p in range(self._np)])

    @property
    def wall(self):
        """The reference's `wall.tiles` Vec: front = rinshan side, back = next live draw."""
        s = self._state()
        return [s.wall[i] for i in range(s.rinshan_draw_count, s.wall_top)]

    @staticmethod
    def _store_wall(s, tiles, front):
        """Vec index i lives at absolute slot front + i: the dora / ura indicators of the reference's
        `(4 + 2k) - rinshan_draw_count` (state/mod.rs:2026,2051) are then the fixed slots 4 + 2k / 5 + 2k."""
        if front + len(tiles) > len(s.wall):
            raise ValueError(f"wall of {len(tiles)} tiles does not fit behind {front} rinshan draws")
        for i in range(len(s.wall)):
            s.wall[i] = 255
        for i, t in enumerate(tiles):
            s.wall[front + i] = int(t)
        s.rinshan_draw_count = front
        s.wall_top = front + len(tiles)

    @wall.setter
    def wall(self, v):                                    # env.rs:139-142 — replaces the Vec, counters untouched
        self._mutate(lambda s: self._store_wall(s, list(v), s.rinshan_draw_count))

    @property
    def rinshan_draw_count(self):
        return self._state().rinshan_draw_count

    @rinshan_draw_count.setter
    def rinshan_draw_count(self, v):                      # env.rs:264-266 — the counter alone; the Vec stays as it is
        self._mutate(lambda s: self._store_wall(s, [s.wall[i] for i in range(s.rinshan_draw_count, s.wall_top)], int(v)))

    @property
    def dora_indicators(self):
        s = self._state()
        return [s.dora_ind[k] for k in range(s.n_dora)]

    @dora_indicators.setter
    def dora_indicators(self, v):                         # env.rs:254-257
        v = [int(t) for t in v]
        if len(v) > 5:
            raise ValueError("at most 5 dora indicators")

        def f(s):
            s.n_dora = len(v)
            for k in range(5):
                s.dora_ind[k] = v[k] if k < len(v) else 255
        self._mutate(f)

    @property
    def last_discard(self):                               # env.rs:560-567: (seat, tile)
        s = self._state()
        return None if s.last_discard_pid == 255 else (s.last_discard_pid, s.last_discard_tile)

    @last_discard.setter
    def last_discard(self, v):
        def f(s):
            s.last_discard_pid, s.last_discard_tile = (255, 255) if v is None else (int(v[0]), int(v[1]))
        self._mutate(f)

    @property
    def current_claims(self):                             # env.rs:551-557: {seat: [Action]} (seats without claims absent)
        s = self._state()
        out = {}
        for p in range(self._np):
            if s.n_claims[p]:
                out[p] = [self._claim_action(s.claims[p][k], p) for k in range(min(s.n_claims[p], A.MAX_CLAIMS))]
        return out

    def _claim_action(self, w, p):
        ty, tile, c = w & 0xFF, (w >> 8) & 0xFF, [(w >> 16) & 0xFF, (w >> 24) & 0xFF]
        cons = [t for t in c if t != 255]
        if ty == ActionType.DAIMINKAN and cons:   # three consumed tiles: the third is the remaining copy of the kind
            kind = cons[0] // 4
            cons = [t for t in range(4 * kind, 4 * kind + 4) if t != tile]
        return self._action_cls(ty, None if tile == 255 else tile, cons, p)

    @current_claims.setter
    def current_claims(self, v):
        def f(s):
            for p in range(4):
                acts = list(v.get(p, ())) if p < self._np else []
                if len(acts) > A.MAX_CLAIMS:
                    raise ValueError(f"at most {A.MAX_CLAIMS} claims per seat")
                s.n_claims[p] = len(acts)
                for k, a in enumerate(acts):
                    c = list(a.consume_tiles[:2]) + [255, 255]
                    s.claims[p][k] = int(a.action_type) | ((255 if a.tile is None else a.tile) << 8) | (c[0] << 16) | (c[1] << 24)
        self._mutate(f)

    @property
    def pao(self):                                        # env.rs:570-583: per seat {yaku id: liable seat}
        s = self._state()
        return [{y: s.pao[p][k] for k, y in enumerate((37, 50)) if s.pao[p][k] != 255} for p in range(self._np)]

    @pao.setter
    def pao(self, v):
        if len(v) != self._np:
            return

        def f(s):
            for p in range(self._np):
                for k, y in enumerate((37, 50)):
                    s.pao[p][k] = int(v[p][y]) if y in v[p] else 255
        self._mutate(f)

    @property
    def kita_tiles(self):                                 # env.rs:331-337 (3P): North tiles set aside, per seat
        out = [[] for _ in range(self._np)]
        if self._np == 3 and not self.skip_mjai_logging:
            for w in self._kyoku_events():
                if w[0] & 0xFF == A.EV_KITA:
                    out[(w[0] >> 16) & 0xFF].append((w[0] >> 24) & 0xFF)
        return out

    @property
    def win_results(self):
        """{seat: WinResult} of the wins declared since the last deal (state/mod.rs:863,1107; cleared by
        _initialize_round, 1729) — rebuilt from the hora events of the binary log."""
        from .hand import WinResult, ordered_yaku

        out = {}
        if self.skip_mjai_logging:
            return out
        s = self._state()
        for w in self._kyoku_events():
            if w[0] & 0xFF != A.EV_HORA:
                continue
            actor, tsumo = (w[0] >> 16) & 0xFF, bool(w[1] & 0xFF)
            han, fu, yakuman = (w[1] >> 16) & 0xFF, (w[1] >> 24) & 0xFF, bool((w[3] >> 8) & 0xFF)
            yaku = ordered_yaku(w[8] | (w[9] << 32))
            pay = (C.c_uint32 * 4)()
            from ._lib import check, lib

            # hand_evaluator.rs:142-155: kazoe scores as 13 han; the evaluator always scores for four seats (3P: three)
            check(lib().rv_calculate_score(13 if (not yakuman and han >= 13) else han, fu, int(actor == s.oya), int(tsumo),
                                           0, self._np, pay))
            pao = next((s.pao[actor][k] for k, y in enumerate((37, 50)) if y in yaku and s.pao[actor][k] != 255), None)
            out[actor] = WinResult(True, yakuman, pay[0], pay[1], pay[2], yaku, han, fu, pao, True)
        return out

    def _kyoku_events(self):
        """events of the binary log since the last start_kyoku, as lists of words"""
        words = self._v.events(0)
        evs, i = [], 0
        while i < len(words):
            n = max(1, (int(words[i]) >> 8) & 0xFF)
            if words[i] & 0xFF == A.EV_START_KYOKU:
                evs = []
            evs.append([int(x) for x in words[i:i + n]])
            i += n
        return evs

    def _reveal_kan_dora(self):                           # env.rs:624-626
        self._v.call(0)

    def _get_ura_markers(self):                           # env.rs:628-630: MJAI strings
        return [tid_to_mjai(t) for t in self._v.call(1)]

    def set_scores(self, pts):                            # env.rs:406-414
        if len(pts) == self._np:
            self._mutate(lambda s: [s.score.__setitem__(p, int(pts[p])) for p in range(self._np)])

    def set_state(self, oya=None, honba=None, riichi_sticks=None, scores=None, round_wind=None):  # env.rs:639-671
        def f(s):
            if oya is not None:
                s.oya = s.kyoku_idx = int(oya)
            if honba is not None:
                s.honba = int(honba)
            if riichi_sticks is not None:
                s.riichi_sticks = int(riichi_sticks)
            if scores is not None and len(scores) == self._np:
                for p in range(self._np):
                    s.score[p] = int(scores[p])
            if round_wind is not None:
                s.round_wind = int(round_wind)
        self._mutate(f)

    def clone(self):                                      # env.rs:359-372: deep copy (record + event log)
        import copy

        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        o = object.__new__(type(self))
        o.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_v"})
        o._event_counts = list(self._event_counts)
        o._log_cache = {k: list(v) for k, v in self._log_cache.items()}
        o._ext_log = None if self._ext_log is None else [list(r) for r in self._ext_log]
        o._v = self._v.clone()
        return o

    __copy__ = clone
