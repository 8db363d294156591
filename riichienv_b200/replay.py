"""Replay ingestion: the reference's `MjaiReplay` / `Kyoku` / step-iterator surface.

Reference: riichienv-core/src/replay/mjai_replay.rs (reader), replay/mod.rs:99-1010 (KyokuStepIterator, KyokuStepIterator3P),
replay/mod.rs:1010-1590 (LogKyoku).  What runs where:

* the log is parsed by the library (`rv_replay_from_jsonl`, csrc/replay.cpp) into fixed records: one `rv_log_kyoku` and a
  list of `rv_log_action` per round;
* the kyoku is tracked by the same game records the simulator uses: `rv_vec_replay_begin` (LogKyoku::steps' set-up) and
  `rv_vec_apply_log_actions` (GameState::apply_log_action) run on the device, and the decision points in between are read
  with the ordinary legal-action / observation calls;
* this file is the host logic of the iterator (which decision a log action stands for, pass observations, the
  reach → discard split), a line-for-line mirror of replay/mod.rs:206-532.

`ReplayBatch` is the batched form the data-parallel path wants: K kyoku replayed in lock-step on one vector of K records.
"""
import ctypes as C

from . import _abi as A
from ._lib import check, lib
from .env import Action, Action3P, ActionType, GameRule, Meld, MeldType, Observation, Observation3P, Phase, RiichiEnv


def _tile_str(t: int) -> str:  # TileConverter::to_string (replay/mod.rs:2224-2241): red fives are "0m" / "0p" / "0s"
    t34 = t // 4
    if t34 // 9 > 3:
        return "?"
    if t in (16, 52, 88):
        return "0" + "mpsz"[t34 // 9]
    return f"{t34 % 9 + 1}{'mpsz'[t34 // 9]}"


class _ActionView:
    """one replay `Action` (replay/mod.rs:35-80) read off an rv_log_action"""

    __slots__ = ("raw", "type", "seat", "tile", "is_liqi", "is_wliqi", "meld_type", "tiles", "froms", "hules", "moqie", "doras",
                 "left_tile_count")

    def __init__(self, a: A.LogAction, aux=None):
        self.raw = a
        # Option fields only a paifu carries (rv_log_action_aux): None for MJAI logs
        self.doras = None if aux is None or aux.n_doras == 0xFF else [aux.doras[i] for i in range(min(aux.n_doras, A.LOG_MAX_DORAS))]
        self.left_tile_count = None if aux is None or aux.left_tile_count == 0xFF else aux.left_tile_count
        self.type, self.seat, self.tile = a.type, a.seat, a.tile
        self.is_liqi, self.is_wliqi = bool(a.flags & 1), bool(a.flags & 2)
        self.moqie = bool(a.flags & 1)
        self.meld_type = a.meld_type
        self.tiles = [a.tiles[k] for k in range(a.n_tiles)]
        self.froms = [a.froms[k] for k in range(a.n_tiles)]
        self.hules = [a.hules[k] for k in range(a.n_hule)]


class LogKyoku:
    """replay/mod.rs:1010-1590 (pyclass `Kyoku`)"""

    def __init__(self, k: A.LogKyoku, actions, rule: GameRule, aux=None):
        self._k = k
        self._actions = actions                  # ctypes array (A.LogAction * n)
        self._aux, self._views_cache = aux, None   # the per-action views are built on first use (steps / events / labels need them)
        self.rule = rule
        n = k.np
        self.scores = [k.scores[p] for p in range(n)]
        self.end_scores = [k.end_scores[p] for p in range(n)]
        self.doras = [k.doras[i] for i in range(min(k.n_doras, A.LOG_MAX_DORAS))]
        self.ura_doras = [k.ura_doras[i] for i in range(min(k.n_ura_doras, A.LOG_MAX_DORAS))]
        self.hands = [[k.hands[p][i] for i in range(k.hand_len[p])] for p in range(n)]
        self.chang, self.ju, self.ben, self.liqibang = k.chang, k.ju, k.ben, k.liqibang
        self.left_tile_count = k.left_tile_count
        self.wliqi = [bool(k.wliqi[p]) for p in range(n)]
        self.paishan = None                      # MJAI logs carry no wall; set by the paifu reader
        self._win_ctx, self._win_error = [], None  # rv_replay_win_contexts of this round (filled by _from_handle)
        self.game_end_scores = [k.game_end_scores[p] for p in range(n)] if k.has_game_end_scores else None

    @property
    def _views(self):
        if self._views_cache is None:
            aux = self._aux
            self._views_cache = [_ActionView(a, aux[i] if aux is not None else None) for i, a in enumerate(self._actions)]
        return self._views_cache

    # ---- features ----------------------------------------------------------------------------------
    def grp_features(self):  # replay/mod.rs:1502-1522
        d = [e - s for s, e in zip(self.scores, self.end_scores)] if len(self.scores) == len(self.end_scores) else []
        return {"chang": self.chang, "ju": self.ju, "ben": self.ben, "liqibang": self.liqibang, "scores": list(self.scores),
                "end_scores": list(self.end_scores), "wliqi": list(self.wliqi), "delta_scores": d}

    def take_grp_features(self):  # replay/mod.rs:1524-1590
        def ranks(scores):  # descending by score, lower seat wins ties
            order = sorted(range(len(scores)), key=lambda i: (-scores[i], i))
            out = [0] * len(scores)
            for r, seat in enumerate(order):
                out[seat] = r
            return out

        init = list(self.scores)
        end = list(self.end_scores) if self.end_scores else init
        ri, re = ranks(init), ranks(end)
        d = {"chang": self.chang, "ju": self.ju, "ben": self.ben, "liqibang": self.liqibang,
             "round_initial_scores": init, "round_end_scores": end, "round_delta_scores": [e - s for s, e in zip(init, end)],
             "round_initial_ranks": ri, "round_end_ranks": list(re), "round_delta_ranks": [e - s for s, e in zip(ri, re)],
             "final_ranks": ranks(self.game_end_scores) if self.game_end_scores is not None else re}
        for i, h in enumerate(self.hands):
            d[f"player{i}_initial_hand_tids"] = list(h)
        return d

    def events(self):  # replay/mod.rs:1294-1500
        data = {"scores": list(self.scores), "doras": [_tile_str(t) for t in self.doras]}
        if self.doras:
            data["dora_marker"] = _tile_str(self.doras[0])
        for i, h in enumerate(self.hands):
            data[f"tiles{i}"] = [_tile_str(t) for t in h]
        data.update(chang=self.chang, ju=self.ju, ben=self.ben, liqibang=self.liqibang, left_tile_count=self.left_tile_count)
        if self.ura_doras:
            data["ura_doras"] = [_tile_str(t) for t in self.ura_doras]
        if self.paishan is not None:
            data["paishan"] = self.paishan
        out = [{"name": "NewRound", "data": data}]
        for a in self._views:
            if a.type == A.LA_DISCARD:
                ev = ("DiscardTile", {"seat": a.seat, "tile": _tile_str(a.tile), "is_liqi": a.is_liqi, "is_wliqi": a.is_wliqi})
                if a.doras is not None:
                    ev[1]["doras"] = [_tile_str(t) for t in a.doras]
            elif a.type == A.LA_DEAL:
                ev = ("DealTile", {"seat": a.seat, "tile": _tile_str(a.tile)})
                if a.doras is not None:
                    ev[1]["doras"] = [_tile_str(t) for t in a.doras]
                if a.left_tile_count is not None:
                    ev[1]["left_tile_count"] = a.left_tile_count
            elif a.type == A.LA_CHI_PENG_GANG:
                mt = {MeldType.Chi: 0, MeldType.Pon: 1, MeldType.Daiminkan: 2, MeldType.Ankan: 3, MeldType.Kakan: 2}[MeldType(a.meld_type)]
                ev = ("ChiPengGang", {"seat": a.seat, "type": mt, "tiles": [_tile_str(t) for t in a.tiles], "froms": list(a.froms)})
            elif a.type == A.LA_ANGANG_ADDGANG:
                d = {"seat": a.seat, "type": 3 if a.meld_type == MeldType.Ankan else 2}
                if a.tiles:
                    d["tiles"] = _tile_str(a.tiles[0])
                ev = ("AnGangAddGang", d)
            elif a.type == A.LA_HULE:
                hs = []
                for h in a.hules:
                    hd = {"seat": h.seat, "hu_tile": _tile_str(h.hu_tile), "zimo": bool(h.zimo), "count": h.count, "fu": h.fu,
                          "fans": [{"id": y} for y in range(64) if (h.fans >> y) & 1], "point_rong": h.point_rong,
                          "point_zimo_qin": h.point_zimo_qin, "point_zimo_xian": h.point_zimo_xian, "yiman": bool(h.yiman)}
                    if h.n_li_doras != 0xFF:
                        hd["li_doras"] = [_tile_str(h.li_doras[i]) for i in range(min(h.n_li_doras, 5))]
                    hs.append(hd)
                ev = ("Hule", {"hules": hs})
            elif a.type == A.LA_DORA:
                ev = ("Dora", {"dora_marker": _tile_str(a.tile)})
            elif a.type == A.LA_BABEI:
                ev = ("BaBei", {"seat": a.seat, "moqie": a.moqie})
            elif a.type == A.LA_NOTILE:
                ev = ("NoTile", {})
            elif a.type == A.LA_LIUJU:
                ev = ("LiuJu", {"type": a.tile, "seat": a.seat, "tiles": [_tile_str(t) for t in a.tiles]})
            else:
                continue
            out.append({"name": ev[0], "data": ev[1]})
        return out

    def take_win_result_contexts(self):  # replay/mod.rs:1089-1091
        return WinResultContextIterator(self)

    # ---- the step iterator -------------------------------------------------------------------------
    def steps(self, seat=None, rule=None, skip_single_action=None):  # replay/mod.rs:1094-1292
        return KyokuStepIterator(self, seat, rule or self.rule, True if skip_single_action is None else bool(skip_single_action))


Kyoku = LogKyoku


class WinResultContext:
    """replay/mod.rs:2097-2180: the hand, melds, indicators and win conditions at one Hule of a log, what the log recorded
    (`expected_*`) and what the evaluator says (`actual`).  The walk that builds it is rv_replay_win_contexts (host); `actual`
    comes from hand_eval_kernel — for all contexts of an iterator / a `verify()` in ONE rv_hand_eval_batch call."""

    def __init__(self, c: A.WinContext, batch):
        q = c.query
        self._c, self._batch = c, batch
        self.seat = c.seat
        self.tiles = [q.tiles[i] for i in range(q.n_tiles)]
        self.melds = []
        for m in range(q.n_melds):
            t = [q.meld_tiles[m][k] for k in range(4) if q.meld_tiles[m][k] != 255]
            called = c.meld_called[m]
            self.melds.append(Meld(q.meld_type[m], t, q.meld_type[m] != MeldType.Ankan, c.meld_from[m], None if called == 255 else called))
        self.agari_tile = q.win_tile
        self.dora_indicators = [q.dora_ind[i] for i in range(q.n_dora)]
        self.ura_indicators = [q.ura_ind[i] for i in range(q.n_ura)]
        from .hand import Conditions

        b = q.cond
        self.conditions = Conditions(*[bool(b >> i & 1) for i in range(9)], player_wind=q.player_wind, round_wind=q.round_wind,
                                     riichi_sticks=0, honba=0, kita_count=q.kita_count)
        self.expected_yaku = [y for y in range(64) if c.expected_yaku >> y & 1]   # ids of HuleData.fans (ascending)
        self.expected_han, self.expected_fu = c.expected_han, c.expected_fu

    @property
    def actual(self):
        return self._batch.result(self)

    def create_calculator(self):
        from .hand import HandEvaluator

        return HandEvaluator(list(self.tiles), list(self.melds))

    def calculate(self, calculator, conditions=None):
        return calculator.calc(self.agari_tile, self.dora_indicators, conditions or self.conditions, self.ura_indicators)

    def __repr__(self):
        return (f"WinResultContext(seat={self.seat}, tiles={self.tiles}, agari_tile={self.agari_tile}, "
                f"expected_han={self.expected_han}, expected_fu={self.expected_fu})")


class _WinBatch:
    """the contexts of one or many kyoku evaluated together: one rv_hand_eval_batch on first use of any `actual`"""

    def __init__(self, ctxs):
        self.ctxs = [WinResultContext(c, self) for c in ctxs]
        self._raw = None

    def raw(self):
        if self._raw is None and self.ctxs:
            from .hand import eval_queries

            self._raw = eval_queries([c._c.query for c in self.ctxs])
        return self._raw or []

    def result(self, ctx):
        from .hand import _to_result

        return _to_result(self.raw()[self.ctxs.index(ctx)])


def _verify_counts(batch):
    """MjSoulReplay::verify (mjsoul_replay.rs:357-430): (wins, wins whose yaku / han / fu differ from the paifu's)"""
    ignored = (1 << 31) | (1 << 32) | (1 << 33)                 # dora, aka dora, ura dora
    yakuman_ids = sum(1 << y for y in range(35, 51))
    mismatches = 0
    for ctx, r in zip(batch.ctxs, batch.raw()):
        exp, sim = ctx._c.expected_yaku, r.yaku_mask
        exp_han = ctx.expected_han * 13 if exp & yakuman_ids and ctx.expected_han < 13 else ctx.expected_han
        bad = (sim & ~ignored) != (exp & ~ignored)
        if not bad:
            want = exp_han - bin(exp & ignored).count("1") + bin(sim & ignored).count("1")
            if exp_han < 13 and r.han != want:
                bad = r.han != exp_han
            elif (r.han >= 13) != (exp_han >= 13):
                bad = True
            if not bad and exp_han < 13 and r.fu != ctx.expected_fu:
                bad = True
        if bad:
            mismatches += 1
            print(f"Mismatch: seat={ctx.seat}, han=(sim={r.han}, exp={ctx.expected_han}), fu=(sim={r.fu}, exp={ctx.expected_fu})")
            print(f"  Expected Yaku: {ctx.expected_yaku}")
            print(f"  Actual Yaku: {ctx.actual.yaku}")
            print(f"  Conditions: {ctx.conditions}")
    return len(batch.ctxs), mismatches


class WinResultContextIterator:
    """replay/mod.rs:1594-1628"""

    def __init__(self, kyoku: LogKyoku):
        if kyoku._win_error:
            raise ValueError(kyoku._win_error)
        self._it = iter(_WinBatch(kyoku._win_ctx).ctxs)

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)


class _ReplayObsEnv:
    """What a replay observation is bound to: the record it was taken from (by value).  The tensor encoders load it into
    the iterator's game vector, run the device encoder and put the live record back."""

    def __init__(self, it, record):
        self._it, self._record = it, record
        self._n_applied = it._idx                # log actions applied when the observation was taken
        self._token = 0
        self._np = it._np
        self.skip_mjai_logging = False
        self._ext_log = None

    def _on_record(self, call):
        v = self._it._env._v
        live = v.get_state(0)
        v.set_state(0, self._record)
        try:
            return call(v)
        finally:
            v.set_state(0, live)

    def _encode(self, pid, extended=False):
        return self._on_record(lambda v: v.encode_single(pid, extended))

    @property
    def _v(self):
        """the vector calls an Observation makes (encode_kawa_single, ...), run on this observation's record"""
        outer = self

        class _OnRecord:
            def __getattr__(self, name):
                return lambda *a, **k: outer._on_record(lambda v: getattr(v, name)(*a, **k))

        return _OnRecord()

    def _encode_seq(self, pid, first_new_event):
        """sparse / numeric / candidates by the device encoder over the record and the event log LogKyoku::steps' set-up left
        (start_kyoku + the dealer's scratch draw: apply_log_action pushes no events — so, as in the reference, numeric takes
        its round-start values from that start_kyoku, and the event-derived `drawn` / `last discarder` see nothing after it);
        progression = the replay-mode cache (rv_replay_progression over the actions applied before this observation)."""
        import numpy as np

        if self._np == 3:
            raise NotImplementedError("the reference has no sequence features for sanma observations")
        sp, nu, _, ca, lens = self._on_record(lambda v: v.encode_seq_single(pid, 0))
        n = self._n_applied
        acts, flags = self._it._kyoku._actions, (C.c_uint8 * max(1, n))(*self._it._tsumogiri[:n])
        m = C.c_int(0)
        check(lib().rv_replay_progression(acts, flags, n, None, 0, C.byref(m)))
        pr = np.zeros((max(1, m.value), 5), np.uint16)
        check(lib().rv_replay_progression(acts, flags, n, pr.ctypes.data_as(C.POINTER(C.c_uint16)), m.value, C.byref(m)))
        lens = lens.copy()
        lens[1] = m.value
        return sp, nu, pr, ca, lens


class KyokuStepIterator:
    """replay/mod.rs:99-532 (4P) and 113-1008 (3P): yields (obs, action) for `seat`, or (seat, obs, action) for seat=None."""

    def __init__(self, kyoku: LogKyoku, seat, rule: GameRule, skip_single_action: bool):
        self._kyoku = kyoku
        self._np = kyoku._k.np
        self._acls = Action3P if self._np == 3 else Action
        # GameState::new(0, false, None, 0, rule): single-round mode, MJAI logging on.  (The reference seeds this scratch state
        # from entropy; its wall only feeds the start_kyoku text of the event log, the logged hands replace the deal.)
        self._env = RiichiEnv(game_mode=3 if self._np == 3 else 0, rule=rule, seed=0)
        arr = (A.LogKyoku * 1)(kyoku._k)
        self._env._v.replay_begin(arr)
        self._env._event_counts = [0, 0, 0, 0]
        self._env._log_cache = {}
        self._env._token += 1
        self._actions = kyoku._views
        self._drawn_seen = (-1, 255)
        self._tsumogiri = [0] * len(kyoku._views)   # per action: the discard was the tile just drawn (the progression cache's moqie)
        self._idx = 0
        self._pending_action = None
        self._filter = seat
        self._skip_single = skip_single_action
        self._pending_pass = []

    def __iter__(self):
        return self

    # ---- state access -----------------------------------------------------------------------------
    def _apply(self, a: _ActionView):
        if a.type == A.LA_DISCARD and self._np == 4:                                   # state/event_handler.rs:343-347
            seen = self._drawn_seen if self._drawn_seen[0] == self._idx else (self._idx, self._env._state().drawn_tile)
            self._tsumogiri[self._idx] = int(seen[1] == a.tile)
        self._env._v.apply_log_actions((A.LogAction * 1)(a.raw))
        self._env._token += 1

    def _observe(self, pid):
        """GameState::get_observation(pid) of the record as it stands, bound by value to that record"""
        env = self._env
        env._token += 1
        s = env._state()
        obs = env._observations([pid], s)[pid]
        obs._env = _ReplayObsEnv(self, s)
        obs._token = 0
        return obs

    def _get_observation_for_replay(self, pid, action, what):  # state/mod.rs:265-328
        env = self._env
        orig = env._state()
        self._drawn_seen = (self._idx, orig.drawn_tile)     # (the record just read: saves _apply a second download)
        if action.action_type in (ActionType.RON, ActionType.CHI, ActionType.PON, ActionType.DAIMINKAN):
            s = env._state()
            s.phase = int(Phase.WaitResponse)
            s.active_mask = 1 << pid
            n = s.n_claims[pid]
            if n < A.MAX_CLAIMS:                      # current_claims.entry(pid).or_default().push(env_action)
                own = [t for t in action.consume_tiles if t != action.tile] or list(action.consume_tiles)
                c = (own + [255, 255])[:2]
                s.claims[pid][n] = int(action.action_type) | ((255 if action.tile is None else action.tile) << 8) | (c[0] << 16) | (c[1] << 24)
                s.n_claims[pid] = n + 1
            env._v.set_state(0, s)
        obs = self._observe(pid)
        if action.action_type == ActionType.KITA:      # 3P: any North tile is equivalent (state_3p/mod.rs:250-257)
            exists = any(a.action_type == ActionType.KITA for a in obs._legal_actions)
        else:
            exists = any(a.action_type == action.action_type and a.tile == action.tile for a in obs._legal_actions)
        keep_riichi_cleared = False
        if not exists and action.action_type == ActionType.DISCARD and (orig.flags[pid] & A.F_RIICHI_DECLARED):
            s = env._state()
            s.flags[pid] &= ~A.F_RIICHI_DECLARED
            env._v.set_state(0, s)
            new_obs = self._observe(pid)
            if any(a.action_type == ActionType.DISCARD and a.tile == action.tile for a in new_obs._legal_actions):
                obs, exists, keep_riichi_cleared = new_obs, True, True
        if keep_riichi_cleared:                       # the reference leaves riichi_declared cleared in this branch
            orig.flags[pid] &= ~A.F_RIICHI_DECLARED
        env._v.set_state(0, orig)                     # phase, active_players, current_claims (and the flag) back
        if not exists:
            raise RuntimeError(f"Replay desync:\n  Env action: {action!r}\n  Log action: {what}\n  Self state:\n"
                               f"    phase: {Phase(orig.phase)!r}\n    drawn: {None if orig.drawn_tile == 255 else orig.drawn_tile}")
        return obs

    def _collect_pass_observations(self, discarder, tile, claimers):  # replay/mod.rs:130-181
        env = self._env
        orig = env._state()
        env._v.call(6)                                # _get_claim_actions_for_player(i, discarder, tile) for every seat
        listed = env._state()
        for i in range(self._np):
            if i == discarder or i in claimers or listed.n_claims[i] == 0:
                continue
            n = min(listed.n_claims[i], A.MAX_CLAIMS)
            had_ron = any((listed.claims[i][k] & 0xFF) == ActionType.RON for k in range(n))
            tmp = A.GameState.from_buffer_copy(bytes(orig))
            tmp.phase = int(Phase.WaitResponse)
            tmp.active_mask = 1 << i
            for p in range(4):
                tmp.n_claims[p] = 0
            tmp.n_claims[i] = n
            for k in range(n):
                tmp.claims[i][k] = listed.claims[i][k]
            env._v.set_state(0, tmp)
            self._pending_pass.append((i, self._observe(i)))
            if had_ron:                               # passing on a ron: same-turn furiten (permanent in riichi)
                orig.flags[i] |= A.F_MISSED_AGARI_DOUJUN
                if orig.flags[i] & A.F_RIICHI_DECLARED:
                    orig.flags[i] |= A.F_MISSED_AGARI_RIICHI
        env._v.set_state(0, orig)

    def _peek_next_claimers(self):  # replay/mod.rs:183-198
        if self._idx >= len(self._actions):
            return []
        a = self._actions[self._idx]
        if a.type == A.LA_CHI_PENG_GANG:
            return [a.seat]
        if a.type == A.LA_HULE:
            return [h.seat for h in a.hules if not h.zimo]
        return []

    def _emit(self, pid, obs, action, single_filter=True):
        """the tail every arm of __next__ shares: filter by seat, drop forced decisions; None = keep iterating"""
        if self._filter is not None:
            if pid != self._filter:
                return None
            if single_filter and self._skip_single and len(obs._legal_actions) <= 1:
                return None
            return (obs, action)
        return (pid, obs, action)

    # ---- __next__ (replay/mod.rs:206-532; the 3P iterator differs in BaBei / Kita handling only) ----
    def __next__(self):
        acts = self._actions
        A_ = self._acls
        while True:
            if self._pending_pass:
                pid, obs = self._pending_pass.pop()
                out = self._emit(pid, obs, A_(ActionType.PASS, None, [], pid))
                if out is not None:
                    return out
                continue
            if self._pending_action is not None:
                pid, action = self._pending_action
                self._pending_action = None
                s = self._env._state()
                staged = False
                if action.action_type == ActionType.DISCARD and not (s.flags[pid] & (A.F_RIICHI_DECLARED | A.F_RIICHI_STAGE)):
                    s.flags[pid] |= A.F_RIICHI_STAGE      # the discard after a reach: legal actions of a declared riichi
                    self._env._v.set_state(0, s)
                    staged = True
                try:
                    obs = self._get_observation_for_replay(pid, action, repr(acts[self._idx].raw.type))
                finally:
                    if staged:
                        s = self._env._state()
                        s.flags[pid] &= ~A.F_RIICHI_STAGE
                        self._env._v.set_state(0, s)
                self._apply(acts[self._idx])
                self._idx += 1
                if action.action_type == ActionType.DISCARD and action.tile is not None:
                    self._collect_pass_observations(pid, action.tile, self._peek_next_claimers())
                out = self._emit(pid, obs, action)
                if out is not None:
                    return out
                continue
            if self._idx >= len(acts):
                raise StopIteration
            a = acts[self._idx]
            if a.type in (A.LA_DEAL, A.LA_DORA, A.LA_NOTILE, A.LA_LIUJU) or (a.type == A.LA_BABEI and self._np == 4):
                self._apply(a)
                self._idx += 1
            elif a.type == A.LA_NONE:
                self._idx += 1
            elif a.type == A.LA_BABEI:                    # 3P: a Kita decision (replay/mod.rs:723-755)
                pid = a.seat
                action = A_(ActionType.KITA, None, [], None)
                obs = self._get_observation_for_replay(pid, action, "BaBei")
                self._apply(a)
                self._idx += 1
                out = self._emit(pid, obs, action)
                if out is not None:
                    return out
            elif a.type == A.LA_DISCARD:
                pid = a.seat
                action = A_(ActionType.DISCARD, a.tile, [], None)
                if a.is_liqi:
                    riichi = A_(ActionType.RIICHI, None, [], None)
                    obs = self._get_observation_for_replay(pid, riichi, "DiscardTile")
                    self._pending_action = (pid, action)
                    out = self._emit(pid, obs, riichi, single_filter=False)
                    if out is not None:
                        return out
                else:
                    obs = self._get_observation_for_replay(pid, action, "DiscardTile")
                    self._apply(a)
                    self._idx += 1
                    self._collect_pass_observations(pid, a.tile, self._peek_next_claimers())
                    out = self._emit(pid, obs, action)
                    if out is not None:
                        return out
            elif a.type == A.LA_CHI_PENG_GANG:
                pid = a.seat
                ty = {MeldType.Chi: ActionType.CHI, MeldType.Pon: ActionType.PON, MeldType.Daiminkan: ActionType.DAIMINKAN}.get(
                    MeldType(a.meld_type), ActionType.CHI)
                action = A_(ty, a.tiles[0] if a.tiles else None, list(a.tiles), None)
                obs = self._get_observation_for_replay(pid, action, "ChiPengGang")
                self._apply(a)
                self._idx += 1
                out = self._emit(pid, obs, action)
                if out is not None:
                    return out
            elif a.type == A.LA_ANGANG_ADDGANG:
                pid = a.seat
                t0 = a.tiles[0] if a.tiles else 0
                if a.meld_type == MeldType.Kakan:
                    cons = []
                    for m in self._env.melds[pid]:
                        if m.meld_type == MeldType.Pon and m.tiles[0] // 4 == t0 // 4:
                            cons = list(m.tiles)
                            break
                    action = A_(ActionType.KAKAN, t0, cons, None)
                else:
                    lo = (t0 // 4) * 4
                    action = A_(ActionType.ANKAN, lo, [lo, lo + 1, lo + 2, lo + 3], None)
                obs = self._get_observation_for_replay(pid, action, "AnGangAddGang")
                self._apply(a)
                self._idx += 1
                out = self._emit(pid, obs, action)
                if out is not None:
                    return out
            elif a.type == A.LA_HULE:
                first = a.hules[0]
                pid = first.seat
                s = self._env._state()
                tsumo = bool(first.zimo) and pid == s.current_player
                if tsumo:
                    action = A_(ActionType.TSUMO, None if s.drawn_tile == 255 else s.drawn_tile, [], None)
                else:
                    action = A_(ActionType.RON, None if s.last_discard_pid == 255 else s.last_discard_tile, [], None)
                obs = self._get_observation_for_replay(pid, action, "Hule")
                self._apply(a)
                self._idx += 1
                out = self._emit(pid, obs, action)
                if out is not None:
                    return out
            else:
                self._idx += 1


KyokuStepIterator3P = KyokuStepIterator


class KyokuIterator:
    """mjai_replay.rs:36-62"""

    def __init__(self, replay):
        self._replay, self._i = replay, 0

    def __iter__(self):
        return self

    def __next__(self):
        if self._i >= len(self._replay.rounds):
            raise StopIteration
        self._i += 1
        return self._replay.rounds[self._i - 1]


class MjaiReplay:
    """mjai_replay.rs:28-385: `MjaiReplay.from_jsonl(path, rule=None)` — plain or gzip JSON lines (detected by content)."""

    def __init__(self, rounds):
        self.rounds = rounds

    @staticmethod
    def _rule(rule):
        if rule is None or rule == "tenhou":
            return GameRule.default_tenhou()
        if rule == "mjsoul":
            return GameRule.default_mjsoul()
        raise ValueError(f"Unknown rule: '{rule}'. Expected 'tenhou' or 'mjsoul'")

    @classmethod
    def _from_handle(cls, h, rule):
        try:
            rounds = []
            for r in range(lib().rv_replay_num_rounds(h)):
                k = A.LogKyoku()
                check(lib().rv_replay_kyoku(h, r, C.byref(k)))
                acts = (A.LogAction * max(1, k.n_actions))()
                n = C.c_int(0)
                check(lib().rv_replay_actions(h, r, acts, k.n_actions, C.byref(n)))
                aux = (A.LogActionAux * max(1, k.n_actions))()
                check(lib().rv_replay_actions_aux(h, r, aux, k.n_actions, C.byref(n)))
                rounds.append(LogKyoku(k, (A.LogAction * k.n_actions).from_buffer_copy(bytes(acts)[: C.sizeof(A.LogAction) * k.n_actions])
                                       if k.n_actions else (A.LogAction * 0)(), rule, aux))
                ky = rounds[-1]
                if lib().rv_replay_win_contexts(h, r, None, 0, C.byref(n)) == 0:
                    ky._win_ctx = (A.WinContext * n.value)()
                    check(lib().rv_replay_win_contexts(h, r, ky._win_ctx, n.value, C.byref(n)))
                else:                        # a log the walk cannot follow: reported when the contexts are asked for
                    ky._win_error = lib().rv_last_error().decode("utf-8", "replace")
                check(lib().rv_replay_paishan(h, r, None, 0, C.byref(n)))
                if n.value >= 0:
                    buf = C.create_string_buffer(n.value + 1)
                    check(lib().rv_replay_paishan(h, r, buf, n.value, C.byref(n)))
                    ky.paishan = buf.raw[: n.value].decode()
            return cls(rounds)
        finally:
            lib().rv_replay_free(h)

    @classmethod
    def from_jsonl(cls, path, rule=None):
        g = cls._rule(rule)
        h = C.c_void_p()
        check(lib().rv_replay_from_jsonl(str(path).encode(), g.bits(), C.byref(h)))
        return cls._from_handle(h, g)

    @classmethod
    def from_text(cls, text, rule=None):
        """the same reader over JSON lines already in memory (not in the reference: its reader only takes a path)"""
        g = cls._rule(rule)
        data = text.encode() if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        check(lib().rv_replay_from_text(data, len(data), g.bits(), C.byref(h)))
        return cls._from_handle(h, g)

    def num_rounds(self):
        return len(self.rounds)

    def take_kyokus(self):
        return KyokuIterator(self)


class MjSoulReplay:
    """replay/mjsoul_replay.rs:20-343: a MjSoul paifu as JSON records {"name": "NewRound" | "DealTile" | ..., "data": {...}} —
    the format `Kyoku.events()` writes.  `from_json(path)` reads a gzip file holding {"rounds": [[...], ...]}; `from_dict`
    takes the paifu dict ({"header", "data"}) or the bare list of rounds and also fills `game_end_scores`."""

    def __init__(self, rounds):
        self.rounds = rounds

    @classmethod
    def from_json(cls, path):
        g = GameRule.default_mjsoul()
        h = C.c_void_p()
        check(lib().rv_replay_from_mjsoul_json(str(path).encode(), g.bits(), C.byref(h)))
        return cls(MjaiReplay._from_handle(h, g).rounds)

    @classmethod
    def from_dict(cls, paifu):
        import json

        g = GameRule.default_mjsoul()
        data = json.dumps(paifu).encode()
        h = C.c_void_p()
        check(lib().rv_replay_from_mjsoul_text(data, len(data), g.bits(), C.byref(h)))
        self = cls(MjaiReplay._from_handle(h, g).rounds)
        if self.rounds:                       # the last round replayed to its end gives the game's final scores (mjsoul_replay.rs:259-340)
            last = self.rounds[-1]
            env = RiichiEnv(game_mode=3 if last._k.np == 3 else 0, rule=g, seed=0)
            env._v.replay_begin((A.LogKyoku * 1)(last._k))
            for a in last._views:
                if a.type != A.LA_NONE:
                    env._v.apply_log_actions((A.LogAction * 1)(a.raw))
            last.end_scores = env.scores()
            for p, v in enumerate(last.end_scores):
                last._k.end_scores[p] = v
            for r in self.rounds:
                r.game_end_scores = list(last.end_scores)
                r._k.has_game_end_scores = 1           # (the records carry them too: ReplayBatch.round_features reads records)
                for p, v in enumerate(last.end_scores):
                    r._k.game_end_scores[p] = v
        return self

    def num_rounds(self):
        return len(self.rounds)

    def take_kyokus(self):
        return KyokuIterator(self)

    @staticmethod
    def verify_files(paths, threads=0):
        """`verify()` over many paifu files at once: parsed by the library's host thread pool (rv_replay_from_files), every win
        of every readable file walked (rv_replay_win_contexts) and evaluated in ONE rv_hand_eval_batch.
        -> (total_agari, total_mismatches, unreadable_files)"""
        g = GameRule.default_mjsoul()
        paths = [str(p).encode() for p in paths]
        arr = (C.c_char_p * max(1, len(paths)))(*paths)
        h, failed, n = C.c_void_p(), C.c_int(0), C.c_int(0)
        check(lib().rv_replay_from_files(arr, len(paths), 1, g.bits(), int(threads), C.byref(h), C.byref(failed)))
        try:
            check(lib().rv_replay_win_contexts(h, -1, None, 0, C.byref(n)))
            ctxs = (A.WinContext * max(1, n.value))()
            check(lib().rv_replay_win_contexts(h, -1, ctxs, n.value, C.byref(n)))
        finally:
            lib().rv_replay_free(h)
        total, bad = _verify_counts(_WinBatch([ctxs[i] for i in range(n.value)]))
        return total, bad, failed.value

    def verify(self):  # mjsoul_replay.rs:357-430 -> (total_agari, total_mismatches); every win of the game in one batch
        for r in self.rounds:
            if r._win_error:
                raise ValueError(r._win_error)
        return _verify_counts(_WinBatch([c for r in self.rounds for c in r._win_ctx]))


class ReplayBatch:
    """K kyoku replayed in lock-step on ONE vector of K game records (the data-parallel form of `Kyoku.steps`): call
    `advance()` until it returns False; after each call `self.vec` holds every kyoku one log action further, and the usual
    batched readers (legal_actions, encode, encode_extended, get_state) see all K positions at once."""

    def __init__(self, kyokus, device=0):
        from .vec_env import VecRiichiEnv

        kyokus = list(kyokus)
        if not kyokus:
            raise ValueError("no kyoku")
        np_ = kyokus[0]._k.np
        if any(k._k.np != np_ for k in kyokus):
            raise ValueError("a batch holds kyoku of one variant (all 4P or all sanma)")
        self.kyokus = kyokus
        self.n = len(kyokus)
        self.vec = VecRiichiEnv(self.n, 3 if np_ == 3 else 0, kyokus[0].rule.bits(), seed_base=0, log_cap_words=0, device=device)
        # the logs go to the device ONCE (rv_vec_replay_load); every advance() is then a kernel launch, nothing crosses PCIe
        first = (C.c_int64 * (self.n + 1))()
        for i, k in enumerate(kyokus):
            first[i + 1] = first[i] + len(k._views)
        total = first[self.n]
        blob = b"".join(bytes(k._actions) for k in kyokus)
        acts = (A.LogAction * max(1, total)).from_buffer_copy(blob.ljust(C.sizeof(A.LogAction) * max(1, total), b"\0"))
        self.vec.replay_load((A.LogKyoku * self.n)(*[k._k for k in kyokus]), acts, first)
        self.position = 0

    @classmethod
    def from_files(cls, paths, rule=None, sanma=False, mjsoul=False, threads=0, device=0, staging=None):
        """The bulk form for a data loader: the files are parsed by a pool of host threads inside the library
        (rv_replay_from_files), the rounds of the wanted variant flattened in C (rv_replay_flatten) and uploaded once — no Python
        object per round or action.  Files that do not parse are skipped and counted in `self.n_failed`, as the reference's
        datasets skip them (riichienv-ml/.../datasets/mjai_logs.py:80-84).  `self.round_index[i]` = position of kyoku i among all
        rounds read (both variants), `self.kyokus` stays empty: labels come from rv_replay_own_turn_labels.
        `staging`: an optional uint8 torch tensor in PINNED host memory that a loader keeps across batches; the action records
        are flattened into it and uploaded from there (the records of a large batch are ~1 GB: from pageable memory the upload
        is the largest part of the load)."""
        import numpy as np

        from .vec_env import VecRiichiEnv

        g = GameRule.default_mjsoul() if (mjsoul and rule is None) else MjaiReplay._rule(rule)
        paths = [str(p).encode() for p in paths]
        arr = (C.c_char_p * max(1, len(paths)))(*paths)
        h, failed = C.c_void_p(), C.c_int(0)
        check(lib().rv_replay_from_files(arr, len(paths), 1 if mjsoul else 0, g.bits(), int(threads), C.byref(h), C.byref(failed)))
        try:
            np_ = 3 if sanma else 4
            nr, na = C.c_int64(0), C.c_int64(0)
            check(lib().rv_replay_totals(h, np_, C.byref(nr), C.byref(na)))
            if nr.value == 0:
                raise ValueError(f"no {'sanma' if sanma else '4-player'} kyoku in {len(paths)} files ({failed.value} unreadable)")
            self = cls.__new__(cls)
            self.kyokus, self.n, self.n_failed, self.position = [], nr.value, failed.value, 0
            ky = (A.LogKyoku * nr.value)()
            # (numpy.empty: the action array of a large batch is 136 B x millions of actions — not worth zeroing first)
            need = max(1, na.value) * C.sizeof(A.LogAction)
            if staging is not None and staging.numel() * staging.element_size() >= need:
                acts = (A.LogAction * max(1, na.value)).from_address(staging.data_ptr())
            else:
                acts_mem = np.empty(need, np.uint8)
                acts = (A.LogAction * max(1, na.value)).from_buffer(acts_mem)
            first = (C.c_int64 * (nr.value + 1))()
            self.round_index = np.zeros(nr.value, np.int32)
            check(lib().rv_replay_flatten(h, np_, ky, acts, first, self.round_index.ctypes.data_as(C.POINTER(C.c_int32))))
        finally:
            lib().rv_replay_free(h)
        self._first = np.frombuffer(first, np.int64).copy()
        seat_flat = np.zeros(max(1, na.value), np.int16)
        id_flat = np.zeros(max(1, na.value), np.int16)
        check(lib().rv_replay_own_turn_labels(acts, na.value, np_, seat_flat.ctypes.data_as(C.POINTER(C.c_int16)),
                                              id_flat.ctypes.data_as(C.POINTER(C.c_int16))))
        self._ky = ky
        self._flat_labels = (seat_flat, id_flat)   # per action, in the order of the flattened logs; labels() makes them [K, T]
        self._labels = None
        self.vec = VecRiichiEnv(self.n, 3 if sanma else 0, g.bits(), seed_base=0, log_cap_words=0, device=device)
        self.vec.replay_load(ky, acts, first)
        return self

    # ---- per-kyoku features for all K rounds at once --------------------------------------------------------------------
    @staticmethod
    def round_features_of(ky, np_):
        """`Kyoku.take_grp_features()` (replay/mod.rs:1524-1590) of every record of an rv_log_kyoku array as numpy arrays:
        chang / ju / ben / liqibang [K], round_initial_scores / round_end_scores / round_delta_scores [K, np],
        round_initial_ranks / round_end_ranks / round_delta_ranks / final_ranks [K, np] (rank 0 = top, the lower seat wins ties)."""
        import numpy as np

        a = np.ctypeslib.as_array(ky)
        out = {k: a[k].astype(np.int32) for k in ("chang", "ju", "ben", "liqibang")}
        init = a["scores"][:, :np_].astype(np.int64)
        end = a["end_scores"][:, :np_].astype(np.int64)

        def ranks(sc):  # descending by score, then by seat
            order = np.argsort(-sc, axis=1, kind="stable")
            r = np.empty_like(order)
            np.put_along_axis(r, order, np.arange(sc.shape[1])[None, :].repeat(sc.shape[0], 0), axis=1)
            return r.astype(np.int32)

        ri, re = ranks(init), ranks(end)
        game_end = a["game_end_scores"][:, :np_].astype(np.int64)
        has_final = a["has_game_end_scores"].astype(bool)
        out.update(round_initial_scores=init.astype(np.int32), round_end_scores=end.astype(np.int32),
                   round_delta_scores=(end - init).astype(np.int32), round_initial_ranks=ri, round_end_ranks=re,
                   round_delta_ranks=re - ri, final_ranks=np.where(has_final[:, None], ranks(game_end), re))
        return out

    def round_features(self):
        if getattr(self, "_ky", None) is None:
            self._ky = (A.LogKyoku * self.n)(*[k._k for k in self.kyokus])
        return self.round_features_of(self._ky, self._ky[0].np)

    # ---- labels: what the seat on turn decided, read off the log -------------------------------------------------------
    @staticmethod
    def own_turn_label(a: _ActionView, np_):
        """(seat, action id) when log action `a` is a decision of the seat on turn — the sample `Kyoku.steps` yields for it
        (replay/mod.rs:206-532; ids: Action::encode, action.rs:158-227, sanma action_3p.rs) — else None.  A riichi discard is
        the Riichi decision (the discard that follows it is taken in a state this walk does not stop at); calls and ron are
        responses to another seat's discard, taken in states `get_observation_for_replay` builds: see `Kyoku.steps`."""
        acls = Action3P if np_ == 3 else Action
        if a.type == A.LA_DISCARD:
            act = acls(ActionType.RIICHI, None, [], None) if a.is_liqi else acls(ActionType.DISCARD, a.tile, [], None)
        elif a.type == A.LA_ANGANG_ADDGANG and a.tiles:
            lo = (a.tiles[0] // 4) * 4
            act = acls(ActionType.ANKAN if a.meld_type == MeldType.Ankan else ActionType.KAKAN, lo, [lo], None)
        elif a.type == A.LA_HULE and a.hules and a.hules[0].zimo:
            return a.hules[0].seat, acls(ActionType.TSUMO, None, [], None).encode()
        elif a.type == A.LA_BABEI and np_ == 3:
            act = acls(ActionType.KITA, None, [], None)
        else:
            return None
        return a.seat, act.encode()

    def labels(self):
        """(seat, action_id): int16 arrays [K, T] (T = the longest log), -1 where log action t of kyoku k is not an own-turn
        decision.  Row (k, seat) of `vec.encode(...)` taken at `position` t is the observation that decision was made on."""
        import numpy as np

        if getattr(self, "_labels", None) is None and getattr(self, "_flat_labels", None) is not None:
            lens = np.diff(self._first)
            T = int(lens.max()) if len(lens) else 0
            seat = np.full((self.n, T), -1, np.int16)
            aid = np.full((self.n, T), -1, np.int16)
            col = np.arange(T)[None, :]
            inside = col < lens[:, None]
            src = (self._first[:-1, None] + col)[inside]
            seat[inside], aid[inside] = self._flat_labels[0][src], self._flat_labels[1][src]
            self._labels = (seat, aid)
        if getattr(self, "_labels", None) is None:
            T = max(len(k._views) for k in self.kyokus)
            seat = np.full((self.n, T), -1, np.int16)
            aid = np.full((self.n, T), -1, np.int16)
            np_ = self.kyokus[0]._k.np
            for i, k in enumerate(self.kyokus):
                for t, a in enumerate(k._views):
                    lab = self.own_turn_label(a, np_)
                    if lab is not None:
                        seat[i, t], aid[i, t] = lab
            self._labels = (seat, aid)
        return self._labels

    def labels_of_rows(self, index, n_rows):
        """for the rows `vec.encode(..., index=index)` just wrote at the current position: the logged action id of each row's
        (game, seat), or -1 when that seat's next logged action is not an own-turn decision (int64 tensor on `index`'s device)"""
        import torch

        t = self.position
        if getattr(self, "_flat_labels", None) is not None:          # from_files: gather from the per-action arrays, no [K, T] table
            if getattr(self, "_flat_dev", None) is None or self._flat_dev[0].device != index.device:
                self._flat_dev = tuple(torch.from_numpy(a).to(index.device) for a in (self._first, *self._flat_labels))
            first, fseat, fid = self._flat_dev
            idx = index[:n_rows].to(torch.int64)
            g, s = idx // 4, idx % 4
            pos = first[g] + t
            ok = pos < first[g + 1]
            pos = torch.where(ok, pos, torch.zeros_like(pos))
            return torch.where(ok & (fseat[pos] == s), fid[pos].to(torch.int64), torch.full_like(g, -1))
        seat, aid = self.labels()
        if t >= seat.shape[1]:
            return torch.full((n_rows,), -1, dtype=torch.int64, device=index.device)
        if getattr(self, "_labels_dev", None) is None or self._labels_dev[0].device != index.device:
            self._labels_dev = (torch.from_numpy(seat.astype("int64")).to(index.device), torch.from_numpy(aid.astype("int64")).to(index.device))
        dseat, daid = self._labels_dev
        idx = index[:n_rows].to(torch.int64)
        g, s = idx // 4, idx % 4
        return torch.where(dseat[g, t] == s, daid[g, t], torch.full_like(g, -1))

    def advance(self):
        """apply the next log action of every kyoku that has one; returns False when every kyoku is exhausted"""
        if self.vec.replay_advance() == 0:
            return False
        self.position += 1
        return True
