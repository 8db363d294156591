"""Uniform single-game driver over the three implementations of the hot path:
  oracle  — oracle/ (CPU restatement of the reference; the checker)
  hostsim — the CUDA device sources compiled for the host (tests/hostsim; CPU debugging aid)
  gpu     — the product: libriichienv_b200.so through its C ABI (VecRiichiEnv of size 1)
Scenario tests are written once against this interface (mirrors tests/env/helper.py of the reference).
"""
import ctypes as C

from riichienv_b200 import _abi as A


def act(type_, tile=None, consume=(), actor=None):
    a = A.Action()
    a.type = type_
    a.tile = 255 if tile is None else tile
    cons = sorted(consume)
    a.n_consume = len(cons)
    for k in range(4):
        a.consume[k] = cons[k] if k < len(cons) else 255
    a.actor = 255 if actor is None else actor
    return a


def act_tuple(a):
    return (a.type, None if a.tile == 255 else a.tile, tuple(a.consume[: a.n_consume]))


class Backend:
    name = "?"

    def step(self, actions):
        arr = (A.Action * 4)()
        for p in range(4):
            arr[p] = actions[p] if p in actions else act(A.NO_ACTION)
        self._step(arr)

    def legal_tuples(self, pid):
        return [act_tuple(a) for a in self.legal(pid)]

    def events_json(self, viewer=-1):
        from riichienv_b200._lib import events_to_json

        return events_to_json(self.events(), viewer)


class OracleBackend(Backend):
    name = "oracle"

    def __init__(self, mode=0, seed=42, rule=A.RULE_DEFAULT_TENHOU):
        import oracle

        self.lib = oracle.load()
        self.h = self.lib.orc_game_new(mode, seed, 0, rule, 1)

    def reset(self, wall=None, oya=0, scores=None, honba=0, kyotaku=0, round_wind=0):
        w = (C.c_uint8 * len(wall))(*wall) if wall is not None else None
        s = (C.c_int32 * 4)(*(list(scores) + [0] * (4 - len(scores)))) if scores is not None else None
        self.lib.orc_game_reset(self.h, oya, round_wind, honba, kyotaku, w, s)

    def get_state(self):
        s = A.GameState()
        self.lib.orc_game_snapshot(self.h, C.byref(s))
        return s

    def set_state(self, s):
        self.lib.orc_game_load_snapshot(self.h, C.byref(s))

    def _step(self, arr):
        self.lib.orc_game_step(self.h, arr)

    def encode_ext(self, pid):
        import numpy as np

        a = np.zeros((215, 34), np.float32)
        self.lib.orc_game_encode_ext(self.h, pid, a.ctypes.data_as(C.POINTER(C.c_float)))
        return a

    def random_step(self, agent_seed, game_id):
        self.lib.orc_game_random_step(self.h, agent_seed, game_id)

    def agent_step(self, policy, agent_seed, game_id):
        self.lib.orc_game_agent_step(self.h, policy, agent_seed, game_id)

    def call(self, op):
        out = (C.c_uint8 * 8)()
        return self.lib.orc_game_call(self.h, op, out)

    def legal(self, pid):
        out = (A.Action * A.MAX_LEGAL)()
        n = self.lib.orc_game_legal(self.h, pid, out)
        return list(out[:n])

    def events_json(self, viewer=-1):
        """the oracle's OWN text log (oracle/json.hpp), not the product's renderer over the oracle's words"""
        n = self.lib.orc_game_mjai_log(self.h, viewer, None, 0)
        buf = C.create_string_buffer(n + 1)
        self.lib.orc_game_mjai_log(self.h, viewer, buf, n + 1)
        return buf.value.decode().split("\n") if n else []

    def events(self):
        n = self.lib.orc_game_events(self.h, None, 0)
        buf = (C.c_uint32 * max(n, 1))()
        self.lib.orc_game_events(self.h, buf, n)
        return list(buf[:n])


class HostsimBackend(Backend):
    name = "hostsim"

    def __init__(self, mode=0, seed=42, rule=A.RULE_DEFAULT_TENHOU):
        from tests import hostsim

        self.lib = hostsim.load()
        self.h = self.lib.hs_game_new(mode, seed, rule, 1 << 16)

    def reset(self, wall=None, oya=0, scores=None, honba=0, kyotaku=0, round_wind=0):
        w = (C.c_uint8 * len(wall))(*wall) if wall is not None else None
        s = (C.c_int32 * 4)(*(list(scores) + [0] * (4 - len(scores)))) if scores is not None else None
        self.lib.hs_game_reset(self.h, oya, round_wind, honba, kyotaku, w, s)

    def get_state(self):
        s = A.GameState()
        self.lib.hs_game_snapshot(self.h, C.byref(s))
        return s

    def set_state(self, s):
        self.lib.hs_game_load_snapshot(self.h, C.byref(s))

    def _step(self, arr):
        self.lib.hs_game_step(self.h, arr)

    def encode_ext(self, pid):
        import numpy as np

        a = np.zeros((215, 34), np.float32)
        self.lib.hs_game_encode_ext(self.h, pid, a.ctypes.data_as(C.POINTER(C.c_float)))
        return a

    def random_step(self, agent_seed, game_id):
        self.lib.hs_game_random_step(self.h, agent_seed, game_id)

    def agent_step(self, policy, agent_seed, game_id):
        self.lib.hs_game_agent_step(self.h, policy, agent_seed, game_id)

    def call(self, op):
        out = (C.c_uint8 * 8)()
        return self.lib.hs_game_call(self.h, op, out)

    def random_step_coopdeal(self, agent_seed, game_id):
        """a random step whose round deal runs through init_round_coop (the warp-cooperative deal of the lock-step kernels)"""
        self.lib.hs_game_random_step_coopdeal(self.h, agent_seed, game_id)

    def visit_deferred(self, agent_seed, game_id):
        """One scheduler visit of the rollout kernels (parked discard tails / deals run on their own visit)."""
        return self.lib.hs_game_random_step_deferred(self.h, agent_seed, game_id)

    def legal(self, pid):
        out = (A.Action * A.MAX_LEGAL)()
        n = self.lib.hs_game_legal(self.h, pid, out)
        return list(out[:n])

    def events(self):
        n = self.lib.hs_game_events(self.h, None, 0)
        buf = (C.c_uint32 * max(n, 1))()
        self.lib.hs_game_events(self.h, buf, n)
        return list(buf[:n])


class GpuBackend(Backend):
    name = "gpu"

    def __init__(self, mode=0, seed=42, rule=A.RULE_DEFAULT_TENHOU):
        from riichienv_b200.vec_env import VecRiichiEnv

        self.v = VecRiichiEnv(1, mode, rule, seeds=[seed], log_cap_words=1 << 16)

    def reset(self, wall=None, oya=0, scores=None, honba=0, kyotaku=0, round_wind=0):
        self.v.reset(oya=oya, round_wind=round_wind, honba=honba, kyotaku=kyotaku,
                     scores=None if scores is None else [scores], walls=None if wall is None else [wall])

    def get_state(self):
        return self.v.get_state(0)

    def set_state(self, s):
        self.v.set_state(0, s)

    def _step(self, arr):
        self.v.step(arr)

    def encode_ext(self, pid):
        """row of seat pid (which must owe an action) through rv_vec_encode_ext"""
        import torch

        obs = torch.zeros((4, 215, 34), dtype=torch.float32, device="cuda")
        idx = torch.full((4,), -1, dtype=torch.int32, device="cuda")
        n = self.v.encode_extended(obs=obs, index=idx, max_obs=4)
        rows = idx[:n].tolist()
        return obs[rows.index(pid)].cpu().numpy()

    def random_step(self, agent_seed, game_id):
        self.v.step_random(agent_seed, 1)

    def agent_step(self, policy, agent_seed, game_id):
        self.v.step_agent(policy, agent_seed, 1)

    def call(self, op):
        return self.v.call(op)

    def legal(self, pid):
        acts, counts = self.v.legal_actions()
        n = int(counts[0, pid])
        return [acts[pid * A.MAX_LEGAL + k] for k in range(n)]

    def events(self):
        return self.v.events(0)


BACKENDS = {"oracle": OracleBackend, "hostsim": HostsimBackend, "gpu": GpuBackend}


def setup_env(backend_cls, seed=42, game_mode=0, hands=None, melds=None, active_players=None, current_player=0,
              phase=0, drawn_tile=None, wall=None, discards=None, riichi_declared=None, points=None, oya=None,
              round_wind=None, rule=A.RULE_DEFAULT_TENHOU):
    """Re-expression of tests/env/helper.py:helper_setup_env on top of snapshot get/set."""
    env = backend_cls(game_mode, seed, rule)
    env.reset(wall=wall, oya=oya or 0)
    s = env.get_state()
    nseats = 3 if game_mode >= 3 else 4

    def set_hand(p, tiles):
        tiles = sorted(tiles)
        for k in range(A.HAND_CAP):
            s.hand[p][k] = tiles[k] if k < len(tiles) else 255
        s.hand_len[p] = len(tiles)

    if hands is not None:
        for p in range(nseats):
            if hands[p] is not None:
                set_hand(p, hands[p])
    if melds is not None:
        for p in range(nseats):
            if melds[p]:
                s.n_melds[p] = len(melds[p])
                for m, (ty, tiles, from_who, called) in enumerate(melds[p]):
                    s.meld_type[p][m] = ty
                    for k in range(4):
                        s.meld_tiles[p][m][k] = tiles[k] if k < len(tiles) else 255
                    s.meld_from[p][m] = 255 if from_who is None or from_who < 0 else from_who
                    s.meld_called[p][m] = 255 if called is None else called
    if active_players is not None:
        s.active_mask = sum(1 << p for p in active_players)
    if current_player is not None:
        s.current_player = current_player
        if active_players is None:
            s.active_mask = 1 << current_player
    if phase is not None:
        s.phase = phase
    s.needs_tsumo = 0
    if drawn_tile is not None:
        s.drawn_tile = drawn_tile
        cur = [s.hand[current_player][k] for k in range(s.hand_len[current_player])]
        set_hand(current_player, cur + [drawn_tile])
    if discards is not None:
        for p in range(nseats):
            s.n_river[p] = len(discards[p])
            s.river_tedashi[p] = (1 << len(discards[p])) - 1
            for k, t in enumerate(discards[p]):
                s.river[p][k] = t
    if riichi_declared is not None:
        for p in range(nseats):
            if riichi_declared[p]:
                s.flags[p] |= A.F_RIICHI_DECLARED
            else:
                s.flags[p] &= ~A.F_RIICHI_DECLARED
    if points is not None:
        for p in range(nseats):
            s.score[p] = points[p]
    if oya is not None:
        s.oya = oya
        s.kyoku_idx = oya
    if round_wind is not None:
        s.round_wind = round_wind
    env.set_state(s)
    return env
