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
    """observation/mod.rs:24-160 — a by-value snapshot for one seat."""

    def __init__(self, env: "RiichiEnv", s: A.GameState, pid: int, legal, new_events, seq_start_word=0):
        self.player_id = pid
        n = env._np
        self.hands = [[s.hand[p][k] for k in range(s.hand_len[p])] if p == pid else [] for p in range(n)]
        self.melds = [_melds_of(s, p) for p in range(n)]
        self.discards = [[s.river[p][k] for k in range(min(s.n_river[p], A.RIVER_CAP))] for p in range(n)]
        self.dora_indicators = [s.dora_ind[k] for k in range(s.n_dora)]
        self.scores = [s.score[p] for p in range(n)]
        self.riichi_declared = [bool(s.flags[p] & A.F_RIICHI_DECLARED) for p in range(n)]
        self._legal_actions = legal
        self._new_events = new_events
        self._env = env
        self._seq_start_word = seq_start_word
        self._token = env._token
        self._seq = None
        self.honba = s.honba
        self.riichi_sticks = s.riichi_sticks
        self.round_wind = s.round_wind
        self.oya = s.oya
        self.kyoku_index = s.kyoku_idx
        self.waits = [k for k in range(34) if (s.c_waits[pid] >> k) & 1]
        self.is_tenpai = bool(self.waits)
        self.tsumogiri_flags = [[], [], [], []]  # always empty in the live env (observation/mod.rs:105)
        self.riichi_sutehais = [None if s.riichi_sutehai[p] == 255 else s.riichi_sutehai[p] for p in range(n)]
        self.last_tedashis = [None if s.last_tedashi[p] == 255 else s.last_tedashi[p] for p in range(n)]
        # state/mod.rs:252 destructures (pid, tile) as (tile, _pid): the field carries the DISCARDER'S SEAT
        self.last_discard = None if s.last_discard_pid == 255 else s.last_discard_pid
        self.drawn_tile = None if s.drawn_tile == 255 else s.drawn_tile

    @property
    def hand(self):
        return list(self.hands[self.player_id])

    @property
    def events(self):
        """Full per-player masked log up to this observation (parsed dicts)."""
        return [json.loads(x) for x in self._env._masked_log(self.player_id)]

    def legal_actions(self):
        return list(self._legal_actions)

    def new_events(self):
        return list(self._new_events)

    @property
    def action_space_size(self):  # observation_3p/python.rs:116-118
        return self._env.action_space_size

    def mask(self):  # observation/python.rs:98-111, observation_3p/python.rs:102-114
        m = bytearray(self._env.action_space_size)
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
        return self._env._encode(self.player_id)

    def encode_extended(self):  # observation/python.rs:1272-1294 (4P)
        return self._ext_row().tobytes()

    def _ext_row(self):
        """(215, 34) float32 array of this observation, computed once by obs_ext_kernel (rv_vec_encode_ext)."""
        import numpy as np

        if getattr(self, "_ext", None) is None:
            if self._token != self._env._token:
                raise RuntimeError("tensors are computed on the device from the live game: call the extended encoders on "
                                   "the observations of the latest reset()/step()")
            self._ext = np.frombuffer(self._env._encode(self.player_id, extended=True), dtype=np.float32).reshape(215, -1)
        return self._ext

    # The standalone encoders of observation/python.rs:195-1270 (sanma: observation_3p/python.rs:198-1115) are the channel
    # blocks of encode_extended — encode.rs holds the same bodies as `_into` variants; a statement-level diff shows only
    # renames — so they are slices of the device-computed row.  Sanma rows carry three relative seats per group and two
    # opponents (NP = 3): the slices stop there.
    def _seats(self):
        return 3 if self._env._np == 3 else 4

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

    def encode_furiten_ron_possibility(self):  # python.rs:251-293 -> (NP, 21)
        """All ones: the encoder only clears a seat's row after three consecutive tsumogiri flags, and the live env never
        fills `tsumogiri_flags` (observation/mod.rs:105) — a constant, so nothing is computed."""
        import numpy as np

        return np.ones((self._seats(), 21), np.float32).tobytes()

    def encode_kawa_overview(self):  # python.rs:881-930 -> (4, 7, 34), seats in absolute order (obs_kawa_kernel; 4P)
        import torch

        if self._token != self._env._token:
            raise RuntimeError("tensors are computed on the device from the live game: call encode_kawa_overview() on the "
                               "observations of the latest reset()/step()")
        env = self._env
        dev = f"cuda:{env._v.ctx.device}"
        out = torch.zeros((4, 4, 7, 34), dtype=torch.float32, device=dev)
        idx = torch.full((4,), -1, dtype=torch.int32, device=dev)
        n = env._v.encode_kawa_overview(out=out, index=idx, max_obs=4)
        rows = idx[:n].tolist()
        if self.player_id not in rows:
            raise ValueError(f"seat {self.player_id} owes no action")
        return out[rows.index(self.player_id)].cpu().numpy().tobytes()

    # ---- sequence features (observation/python.rs:1297-1364): raw bytes, as the reference returns them ----
    def _seq_features(self):
        if self._seq is None:
            if self._token != self._env._token:
                raise RuntimeError("sequence features are computed on the device from the live game: call them on the "
                                   "observations of the latest reset()/step()")
            self._seq = self._env._encode_seq(self.player_id, self._seq_start_word)
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


Observation3P = Observation   # observation_3p/mod.rs:19-46: the same snapshot over 3 seats (lists are sized by the env's seat count)


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
        self._token = getattr(self, "_token", 0) + 1
        self._v.reset(oya=0 if oya is None else oya, round_wind=0 if round_wind is None else round_wind,
                      honba=0 if honba is None else honba, kyotaku=0 if kyotaku is None else kyotaku,
                      scores=None if scores is None else [list(scores)], walls=None if wall is None else [list(wall)])
        return self._observations(self.active_players)

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

    def _get_legal_actions(self, pid: int):
        acts, counts = self._v.legal_actions()
        return [self._action_cls._from_abi(acts[pid * A.MAX_LEGAL + k]) for k in range(int(counts[0, pid]))]

    # ---- logs -----------------------------------------------------------------------------------
    @property
    def mjai_log(self):
        if self.skip_mjai_logging:
            return []
        return [json.loads(x) for x in self._v.mjai_log(0)]

    def _masked_log(self, pid):
        if self.skip_mjai_logging:
            return []
        return events_to_json(self._v.events(0), pid)

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
            start_word = self._event_word_offset(self._event_counts[p])
            self._event_counts[p] = len(full)  # state/mod.rs:211-218: the delta advances on every observation
            out[p] = Observation(self, s, p, legal, new, start_word)
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

    def _encode_seq(self, pid, start_word):
        import numpy as np
        import torch

        if self.skip_mjai_logging:
            raise ValueError("sequence features need the MJAI log (skip_mjai_logging=False)")
        dev = f"cuda:{self._v.ctx.device}"
        sp = torch.zeros((4, 25), dtype=torch.uint16, device=dev)
        nu = torch.zeros((4, 12), dtype=torch.float32, device=dev)
        pr = torch.zeros((4, 512, 5), dtype=torch.uint16, device=dev)
        ca = torch.zeros((4, 64, 4), dtype=torch.uint16, device=dev)
        le = torch.zeros((4, 3), dtype=torch.uint16, device=dev)
        idx = torch.full((4,), -1, dtype=torch.int32, device=dev)
        start = np.zeros((1, 4), np.uint32)
        start[0, pid] = start_word
        n = self._v.encode_seq(sparse=sp, numeric=nu, prog=pr, cand=ca, lens=le, index=idx, game_style=1, max_obs=4, start_words=start)
        rows = idx[:n].tolist()
        if pid not in rows:
            raise ValueError(f"seat {pid} owes no action")
        r = rows.index(pid)
        lens = le[r].cpu().numpy()
        return sp[r].cpu().numpy(), nu[r].cpu().numpy(), pr[r].cpu().numpy(), ca[r].cpu().numpy(), lens

    def _encode(self, pid, extended=False):
        """bytes of the (74, 34) — sanma (74, 27) — float32 tensor for seat `pid` (must owe an action), computed by
        obs_encode_kernel; extended=True: the (215, 34) tensor of encode_extended (obs_ext_kernel, 4P)."""
        import torch

        dev = f"cuda:{self._v.ctx.device}"
        idx = torch.full((4,), -1, dtype=torch.int32, device=dev)
        if extended:
            obs = torch.zeros((4, 215, 27 if self._np == 3 else 34), dtype=torch.float32, device=dev)
            n = self._v.encode_extended(obs=obs, index=idx, max_obs=4)
        else:
            obs = torch.zeros((4, 74, 27 if self._np == 3 else 34), dtype=torch.float32, device=dev)
            n = self._v.encode(obs=obs, index=idx, max_obs=4)
        rows = idx[:n].tolist()
        if pid not in rows:
            raise ValueError(f"seat {pid} owes no action; encode() is defined for the observations step()/reset() return")
        return obs[rows.index(pid)].cpu().numpy().tobytes()

    # ---- getters / setters used by callers and by the reference's tests (env.rs:134-635) --------------
    kyoku_idx = property(lambda self: self._state().kyoku_idx)
    round_wind = property(lambda self: self._state().round_wind)
    oya = property(lambda self: self._state().oya)
    honba = property(lambda self: self._state().honba)
    riichi_sticks = property(lambda self: self._state().riichi_sticks)
    turn_count = property(lambda self: self._state().turn_count)
    is_done = property(lambda self: bool(self._state().is_done))
    num_players = property(lambda self: self._np)
    action_space_size = property(lambda self: 60 if self._np == 3 else 82)
    _action_cls = property(lambda self: Action3P if self._np == 3 else Action)
    last_error = property(lambda self: None if self._state().last_error == 255 else
                          f"Error: Illegal Action by Player {self._state().last_error}")
    dora_indicators = property(lambda self: [self._state().dora_ind[k] for k in range(self._state().n_dora)])

    def _mutate(self, fn):
        s = self._state()
        fn(s)
        self._v.set_state(0, s)

    @property
    def phase(self):
        return Phase(self._state().phase)

    @phase.setter
    def phase(self, v):
        self._mutate(lambda s: setattr(s, "phase", int(v)))

    @property
    def current_player(self):
        return self._state().current_player

    @current_player.setter
    def current_player(self, v):
        self._mutate(lambda s: setattr(s, "current_player", int(v)))

    @property
    def active_players(self):
        return self._active(self._state())

    @active_players.setter
    def active_players(self, v):
        self._mutate(lambda s: setattr(s, "active_mask", sum(1 << int(p) for p in v)))

    @property
    def needs_tsumo(self):
        return bool(self._state().needs_tsumo)

    @needs_tsumo.setter
    def needs_tsumo(self, v):
        self._mutate(lambda s: setattr(s, "needs_tsumo", int(bool(v))))

    @property
    def drawn_tile(self):
        t = self._state().drawn_tile
        return None if t == 255 else t

    @drawn_tile.setter
    def drawn_tile(self, v):
        self._mutate(lambda s: setattr(s, "drawn_tile", 255 if v is None else int(v)))

    @property
    def hands(self):
        s = self._state()
        return [[s.hand[p][k] for k in range(s.hand_len[p])] for p in range(self._np)]

    @hands.setter
    def hands(self, v):
        def f(s):
            for p in range(self._np):
                tiles = list(v[p])[: A.HAND_CAP]
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
        def f(s):
            for p in range(self._np):
                ms = list(v[p])[:4]
                s.n_melds[p] = len(ms)
                for m in range(self._np):
                    for k in range(self._np):
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
        def f(s):
            for p in range(self._np):
                d = list(v[p])[: A.RIVER_CAP]
                s.n_river[p] = len(d)
                s.river_tedashi[p] = (1 << len(d)) - 1
                for k in range(A.RIVER_CAP):
                    s.river[p][k] = d[k] if k < len(d) else 255
        self._mutate(f)

    @property
    def riichi_declared(self):
        s = self._state()
        return [bool(s.flags[p] & A.F_RIICHI_DECLARED) for p in range(self._np)]

    @riichi_declared.setter
    def riichi_declared(self, v):
        def f(s):
            for p in range(self._np):
                s.flags[p] = (s.flags[p] | A.F_RIICHI_DECLARED) if v[p] else (s.flags[p] & ~A.F_RIICHI_DECLARED)
        self._mutate(f)

    @property
    def wall(self):
        """The reference's `wall.tiles` Vec: front = rinshan side, back = next live draw."""
        s = self._state()
        return [s.wall[i] for i in range(s.rinshan_draw_count, s.wall_top)]

    def set_scores(self, pts):
        self._mutate(lambda s: [s.score.__setitem__(p, int(pts[p])) for p in range(self._np)])

    def set_state(self, oya=None, round_wind=None, honba=None, kyotaku=None, scores=None):  # env.rs:636-671
        def f(s):
            if oya is not None:
                s.oya = s.kyoku_idx = int(oya)
            if round_wind is not None:
                s.round_wind = int(round_wind)
            if honba is not None:
                s.honba = int(honba)
            if kyotaku is not None:
                s.riichi_sticks = int(kyotaku)
            if scores is not None and len(scores) == self._np:
                for p in range(self._np):
                    s.score[p] = int(scores[p])
        self._mutate(f)
