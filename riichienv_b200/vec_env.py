"""VecRiichiEnv — N independent Riichi games resident in HBM, advanced by CUDA kernels.

Host-side mirror of a *list* of the reference's `RiichiEnv` objects
(riichienv-ml/src/riichienv_ml/.../_ppo_worker.py:39) behind one handle.  All compute
happens in libriichienv_b200.so; this file only marshals numpy buffers.
"""
import ctypes as C

import numpy as np

from . import _abi as A
from ._lib import Context, check, events_to_json, lib

GAME_MODES = {"4p-red-single": 0, "4p-red-east": 1, "4p-red-half": 2, "3p-red-single": 3, "3p-red-east": 4, "3p-red-half": 5}


def _ptr(a, ty):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


class VecRiichiEnv:
    def __init__(self, n, game_mode="4p-red-half", rule_bits=A.RULE_DEFAULT_TENHOU, seeds=None, seed_base=0,
                 log_cap_words=0, device=0):
        if isinstance(game_mode, str):
            game_mode = GAME_MODES.get(game_mode, 0)  # unknown string -> mode 0 (env.rs:100)
        self.ctx = Context.get(device)
        self.n = int(n)
        self.game_mode = int(game_mode)
        self.handle = C.c_void_p()
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        check(lib().rv_vec_create(self.ctx.handle, self.n, self.game_mode, int(rule_bits), _ptr(s, C.c_uint64),
                                  int(seed_base), int(log_cap_words), C.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().rv_vec_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # RiichiEnv.reset (env.rs:799-851) for all games
    def reset(self, oya=None, round_wind=None, honba=None, kyotaku=None, scores=None, walls=None):
        def arr(x, dt, shape):
            if x is None:
                return None
            a = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=dt), shape))
            return a
        n = self.n
        nplayers = 3 if self.game_mode >= 3 else 4
        wall_len = 108 if self.game_mode >= 3 else 136
        if scores is not None:
            sc = np.asarray(scores, dtype=np.int32)
            if sc.shape[-1] != nplayers:
                raise ValueError(f"scores length {sc.shape[-1]} does not match number of players {nplayers}")
            sc = np.broadcast_to(sc, (n, nplayers))
            scores = np.zeros((n, A.NP), np.int32)   # the C ABI strides scores by 4 seats
            scores[:, :nplayers] = sc
        a_oya, a_rw, a_hb = arr(oya, np.uint8, (n,)), arr(round_wind, np.uint8, (n,)), arr(honba, np.uint8, (n,))
        a_ky, a_sc, a_w = arr(kyotaku, np.uint32, (n,)), arr(scores, np.int32, (n, A.NP)), arr(walls, np.uint8, (n, wall_len))
        check(lib().rv_vec_reset(self.handle, _ptr(a_oya, C.c_uint8), _ptr(a_rw, C.c_uint8), _ptr(a_hb, C.c_uint8),
                                 _ptr(a_ky, C.c_uint32), _ptr(a_sc, C.c_int32), _ptr(a_w, C.c_uint8)))

    def reseed(self, seeds=None, seed_base=0):
        s = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64)
        check(lib().rv_vec_reseed(self.handle, _ptr(s, C.c_uint64), int(seed_base)))

    def step_random(self, agent_seed, max_steps=1):
        done = C.c_uint64(0)
        check(lib().rv_vec_step_random(self.handle, int(agent_seed), int(max_steps), C.byref(done)))
        return int(done.value)

    def step_agent(self, policy, agent_seed, max_steps=1):
        """up to max_steps env steps per game with on-device agent `policy` (0 = uniform random, 1 = greedy-win)"""
        done = C.c_uint64(0)
        check(lib().rv_vec_step_agent(self.handle, int(policy), int(agent_seed), int(max_steps), C.byref(done)))
        return int(done.value)

    def step_random_async(self, agent_seed, max_steps):
        check(lib().rv_vec_step_random_async(self.handle, int(agent_seed), int(max_steps)))

    def steps_total(self):
        s, g = C.c_uint64(0), C.c_int64(0)
        check(lib().rv_vec_steps_total(self.handle, C.byref(s), C.byref(g)))
        return int(s.value), int(g.value)

    def step(self, actions):
        """actions: ctypes array (A.Action * (n*NP)) or numpy structured bytes of the same layout."""
        check(lib().rv_vec_step(self.handle, C.cast(actions, C.POINTER(A.Action))))

    def legal_actions(self):
        acts = (A.Action * (self.n * A.NP * A.MAX_LEGAL))()
        counts = np.zeros((self.n, A.NP), np.uint8)
        check(lib().rv_vec_legal_actions(self.handle, acts, _ptr(counts, C.c_uint8)))
        return acts, counts

    def encode(self, obs=None, mask=None, index=None, max_obs=None, sync=True):
        """Write FEATURE_ENCODING tensors / 82-id masks of every seat that owes an action into DEVICE buffers
        (torch tensors): obs [max_obs,74,34] f32, mask [max_obs,82] u8, index [max_obs] i32.  Returns the row count."""
        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        if max_obs is None:
            max_obs = min(t.shape[0] for t in (obs, mask, index) if t is not None)
        n = C.c_int64(0)
        check(lib().rv_vec_encode(self.handle, ptr(obs), ptr(mask), ptr(index), int(max_obs), C.byref(n) if sync else None))
        return int(n.value) if sync else None

    def encode_extended(self, obs=None, mask=None, index=None, max_obs=None, sync=True):
        """Observation.encode_extended rows (observation/python.rs:1272-1294) of every seat that owes an action, into DEVICE
        buffers: obs [max_obs,215,34] f32, mask [max_obs,82] u8, index [max_obs] i32 (sanma: Observation3P.encode_extended,
        [max_obs,215,27] and 60 mask ids).  Returns the row count."""
        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        if max_obs is None:
            max_obs = min(t.shape[0] for t in (obs, mask, index) if t is not None)
        n = C.c_int64(0)
        check(lib().rv_vec_encode_ext(self.handle, ptr(obs), ptr(mask), ptr(index), int(max_obs), C.byref(n) if sync else None))
        return int(n.value) if sync else None

    def encode_kawa_overview(self, out=None, index=None, max_obs=None, sync=True):
        """Observation.encode_kawa_overview rows (observation/python.rs:881-930) of every seat that owes an action, into DEVICE
        buffers: out [max_obs,4,7,34] f32, index [max_obs] i32.  4P only.  Returns the row count."""
        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        if max_obs is None:
            max_obs = min(t.shape[0] for t in (out, index) if t is not None)
        n = C.c_int64(0)
        check(lib().rv_vec_encode_kawa(self.handle, ptr(out), ptr(index), int(max_obs), C.byref(n) if sync else None))
        return int(n.value) if sync else None

    def observe_step_random(self, agent_seed, obs=None, mask=None, index=None, max_obs=None, sync=False):
        """encode(obs, mask, index) of the current decision point, then one env step of every live game with the on-device
        agent — one fused kernel (rv_vec_observe_step_random).  Returns the row count when sync=True."""
        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        if max_obs is None:
            max_obs = min(t.shape[0] for t in (obs, mask, index) if t is not None)
        n = C.c_int64(0)
        check(lib().rv_vec_observe_step_random(self.handle, int(agent_seed), ptr(obs), ptr(mask), ptr(index), int(max_obs),
                                               C.byref(n) if sync else None))
        return int(n.value) if sync else None

    def encode_seq(self, sparse=None, numeric=None, prog=None, cand=None, lens=None, index=None, game_style=1, max_obs=None,
                   start_words=None, sync=True):
        """Sequence features (Observation.encode_seq_sparse/numeric/progression/candidates) of every seat that owes an
        action, into DEVICE buffers (torch tensors): sparse [max_obs,25] u16 (pad 441), numeric [max_obs,12] f32,
        prog [max_obs,max_prog,5] u16 (pad 4,276,2,2,4), cand [max_obs,64,4] u16 (pad 279,2,2,3), lens [max_obs,3] u16,
        index [max_obs] i32.  The features cover each seat's event delta since its previous observation: by default the
        library's per-seat cursors are used and advanced; `start_words` (numpy u32 [n,4]) overrides them."""
        def ptr(t):
            return None if t is None else C.c_void_p(t.data_ptr())
        if max_obs is None:
            max_obs = min(t.shape[0] for t in (sparse, numeric, prog, cand, lens, index) if t is not None)
        max_prog = int(prog.shape[1]) if prog is not None else 0
        sw = None
        if start_words is not None:
            start_words = np.ascontiguousarray(start_words, np.uint32).reshape(self.n, A.NP)
            sw = _ptr(start_words, C.c_uint32)
        n = C.c_int64(0)
        check(lib().rv_vec_encode_seq(self.handle, int(game_style), sw, ptr(sparse), ptr(numeric), ptr(prog), max_prog, ptr(cand),
                                      ptr(lens), ptr(index), int(max_obs), C.byref(n) if (sync or sw is not None) else None))
        return int(n.value) if (sync or sw is not None) else None

    def results(self):
        done = np.zeros(self.n, np.uint8)
        scores = np.zeros((self.n, A.NP), np.int32)
        ranks = np.zeros((self.n, A.NP), np.uint8)
        check(lib().rv_vec_results(self.handle, _ptr(done, C.c_uint8), _ptr(scores, C.c_int32), _ptr(ranks, C.c_uint8)))
        return done, scores, ranks

    def counters(self):
        sc = np.zeros(self.n, np.uint32)
        kc = np.zeros(self.n, np.uint32)
        ec = np.zeros(self.n, np.uint32)
        eh = np.zeros(self.n, np.uint64)
        check(lib().rv_vec_counters(self.handle, _ptr(sc, C.c_uint32), _ptr(kc, C.c_uint32), _ptr(ec, C.c_uint32),
                                    _ptr(eh, C.c_uint64)))
        return sc, kc, ec, eh

    def get_state(self, game=0):
        s = A.GameState()
        check(lib().rv_vec_get_state(self.handle, int(game), C.byref(s)))
        return s

    def set_state(self, game, state):
        check(lib().rv_vec_set_state(self.handle, int(game), C.byref(state)))

    def apply_events(self, events):
        """GameState::apply_mjai_event for every game: `events` = ctypes array (A.MjaiEvent * n), type 0 = no event"""
        check(lib().rv_vec_apply_events(self.handle, events))

    def replay_begin(self, kyokus):
        """LogKyoku::steps' state set-up for every game: `kyokus` = ctypes array (A.LogKyoku * n)"""
        check(lib().rv_vec_replay_begin(self.handle, kyokus))

    def replay_load(self, kyokus, actions, first):
        """rv_vec_replay_begin + one upload of every record's action list (first = n + 1 offsets into `actions`)"""
        check(lib().rv_vec_replay_load(self.handle, kyokus, actions, first))

    def replay_advance(self, sync=True):
        """the next log action of every record that has one left, from the log in HBM; returns how many records did"""
        n = C.c_int64(0)
        check(lib().rv_vec_replay_advance(self.handle, C.byref(n) if sync else None))
        return int(n.value) if sync else None

    def apply_log_actions(self, actions):
        """GameState::apply_log_action for every game: `actions` = ctypes array (A.LogAction * n), type 0 = no action"""
        check(lib().rv_vec_apply_log_actions(self.handle, actions))

    def clone(self):
        """independent copy of every game (RiichiEnv.clone, env.rs:358-372)"""
        o = object.__new__(type(self))
        o.ctx, o.n, o.game_mode = self.ctx, self.n, self.game_mode
        o.handle = C.c_void_p()
        check(lib().rv_vec_clone(self.handle, C.byref(o.handle)))
        return o

    def call(self, op, game=0):
        """env.rs:624-631 hooks: op 0 reveal_kan_dora -> indicator count; op 1 -> list of ura indicator tile ids;
        ops 2-5 (tests.rs): _trigger_ryukyoku("exhaustive_draw") / _initialize_next_round variants -> is_done"""
        out = (C.c_uint8 * 5)()
        n = C.c_int(0)
        check(lib().rv_vec_debug_call(self.handle, int(game), int(op), out, C.byref(n)))
        return list(out[: n.value]) if op == 1 else n.value

    def state_device_ptr(self):
        p = C.c_void_p()
        check(lib().rv_vec_state_device_ptr(self.handle, C.byref(p)))
        return int(p.value)

    def events(self, game=0):
        n = C.c_uint32(0)
        check(lib().rv_vec_events(self.handle, int(game), None, 0, C.byref(n)))
        buf = (C.c_uint32 * max(1, n.value))()
        check(lib().rv_vec_events(self.handle, int(game), buf, n.value, C.byref(n)))
        return list(buf[: n.value])

    def mjai_log(self, game=0, viewer=-1, skip_events=0):
        """MJAI JSON lines of one game's log (viewer -1: all-seeing; 0..3: that seat's masked view), from event
        `skip_events` on"""
        words = self.events(game)
        i = 0
        for _ in range(skip_events):
            if i >= len(words):
                break
            i += max(1, (int(words[i]) >> 8) & 0xFF)
        return events_to_json(words[i:], viewer)

    # ---- one-row conveniences for the single-env shim (a RiichiEnv is a vector of one game) ----
    def _row_of(self, idx, n, pid):
        rows = idx[:n].tolist()
        if pid not in rows:
            raise ValueError(f"seat {pid} owes no action; the tensors are defined for the observations step()/reset() return")
        return rows.index(pid)

    def encode_single(self, pid, extended=False):
        """bytes of the (74, W) / extended (215, W) float32 tensor of seat `pid` of game 0 (W = 34, sanma 27)."""
        import torch

        dev = f"cuda:{self.ctx.device}"
        w = 27 if self.game_mode >= 3 else 34
        idx = torch.full((4,), -1, dtype=torch.int32, device=dev)
        obs = torch.zeros((4, 215 if extended else 74, w), dtype=torch.float32, device=dev)
        n = (self.encode_extended if extended else self.encode)(obs=obs, index=idx, max_obs=4)
        return obs[self._row_of(idx, n, pid)].cpu().numpy().tobytes()

    def encode_kawa_single(self, pid):
        import torch

        dev = f"cuda:{self.ctx.device}"
        shape = (4, 3, 7, 27) if self.game_mode >= 3 else (4, 4, 7, 34)
        out = torch.zeros(shape, dtype=torch.float32, device=dev)
        idx = torch.full((4,), -1, dtype=torch.int32, device=dev)
        n = self.encode_kawa_overview(out=out, index=idx, max_obs=4)
        return out[self._row_of(idx, n, pid)].cpu().numpy().tobytes()

    def encode_seq_single(self, pid, start_word):
        """(sparse[25] u16, numeric[12] f32, prog[512,5] u16, cand[64,4] u16, lens[3]) numpy arrays of seat `pid`."""
        import torch

        dev = f"cuda:{self.ctx.device}"
        sp = torch.zeros((4, 25), dtype=torch.uint16, device=dev)
        nu = torch.zeros((4, 12), dtype=torch.float32, device=dev)
        pr = torch.zeros((4, 512, 5), dtype=torch.uint16, device=dev)
        ca = torch.zeros((4, 64, 4), dtype=torch.uint16, device=dev)
        le = torch.zeros((4, 3), dtype=torch.uint16, device=dev)
        idx = torch.full((4,), -1, dtype=torch.int32, device=dev)
        start = np.zeros((1, 4), np.uint32)
        start[0, pid] = start_word
        n = self.encode_seq(sparse=sp, numeric=nu, prog=pr, cand=ca, lens=le, index=idx, game_style=1, max_obs=4, start_words=start)
        r = self._row_of(idx, n, pid)
        return sp[r].cpu().numpy(), nu[r].cpu().numpy(), pr[r].cpu().numpy(), ca[r].cpu().numpy(), le[r].cpu().numpy()


class MultiVecRiichiEnv:
    """N games sharded over several GPUs behind one handle (rv_multi_*): device k owns the contiguous global game ids
    [k*N/G, (k+1)*N/G), game g is seeded seed_base + g whatever G is, one host thread and stream per device, no traffic between
    devices on the step path; `stats()` is the end-of-run reduction.  What the reference does with a list of RiichiEnv per Ray
    actor (riichienv-ml/src/riichienv_ml/.../_ppo_worker.py:13,39)."""

    def __init__(self, n, game_mode="4p-red-half", rule_bits=A.RULE_DEFAULT_TENHOU, seed_base=0, log_cap_words=0, devices=(0,)):
        if isinstance(game_mode, str):
            game_mode = GAME_MODES.get(game_mode, 0)
        self.n, self.game_mode, self.devices = int(n), int(game_mode), [int(d) for d in devices]
        self.handle = C.c_void_p()
        dev = (C.c_int * len(self.devices))(*self.devices)
        check(lib().rv_multi_create(dev, len(self.devices), self.n, self.game_mode, int(rule_bits), int(seed_base), int(log_cap_words),
                                    C.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().rv_multi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shard(self, k):
        """(VecRiichiEnv view of device k's games, first global id, count) — for the per-device calls (encoders, snapshots)"""
        vec, first, cnt = C.c_void_p(), C.c_int64(0), C.c_int64(0)
        check(lib().rv_multi_shard(self.handle, int(k), C.byref(vec), C.byref(first), C.byref(cnt)))
        v = object.__new__(VecRiichiEnv)
        v.ctx, v.n, v.game_mode, v.handle = None, int(cnt.value), self.game_mode, vec
        v.close = lambda: None            # owned by the multi-device handle
        return v, int(first.value), int(cnt.value)

    def reset(self):
        check(lib().rv_multi_reset(self.handle))

    def reseed(self, seed_base=0):
        check(lib().rv_multi_reseed(self.handle, int(seed_base)))

    def step_random(self, agent_seed, max_steps=1):
        done = C.c_uint64(0)
        check(lib().rv_multi_step_random(self.handle, int(agent_seed), int(max_steps), C.byref(done)))
        return int(done.value)

    def results(self):
        done = np.zeros(self.n, np.uint8)
        scores = np.zeros((self.n, A.NP), np.int32)
        ranks = np.zeros((self.n, A.NP), np.uint8)
        check(lib().rv_multi_results(self.handle, _ptr(done, C.c_uint8), _ptr(scores, C.c_int32), _ptr(ranks, C.c_uint8)))
        return done, scores, ranks

    def counters(self):
        sc, kc, ec = np.zeros(self.n, np.uint32), np.zeros(self.n, np.uint32), np.zeros(self.n, np.uint32)
        eh = np.zeros(self.n, np.uint64)
        check(lib().rv_multi_counters(self.handle, _ptr(sc, C.c_uint32), _ptr(kc, C.c_uint32), _ptr(ec, C.c_uint32), _ptr(eh, C.c_uint64)))
        return sc, kc, ec, eh

    def stats(self):
        s = A.RunStats()
        check(lib().rv_multi_stats(self.handle, C.byref(s)))
        return {"games": s.games, "games_done": s.games_done, "env_steps": s.env_steps, "rounds": s.rounds,
                "score_sum": list(s.score_sum), "rank_hist": [list(r) for r in s.rank_hist]}
