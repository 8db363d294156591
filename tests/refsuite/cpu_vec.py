"""TEST INFRASTRUCTURE ONLY.  A one-game stand-in for riichienv_b200.vec_env.VecRiichiEnv over the CPU checkers:
  oracle  — oracle/ (the C++ restatement of the reference)
  hostsim — the CUDA device sources compiled for the host (tests/hostsim)
so that the Python shim (riichienv_b200/env.py: setters, observations, logs) and the reference's own pytest suite can
run on the GPU-less authoring box.  The product never imports this; tests/refsuite/riichienv selects it with
RV_REFSUITE_BACKEND=oracle|hostsim (default on a box without a GPU) and uses the CUDA library otherwise.
"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

from riichienv_b200 import _abi as A

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _hostsim():
    """tests/hostsim by file path: while the reference's suite runs, the name `tests` is the reference's own package"""
    if "rv_hostsim" not in sys.modules:
        spec = importlib.util.spec_from_file_location("rv_hostsim", os.path.join(_ROOT, "tests", "hostsim", "__init__.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["rv_hostsim"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["rv_hostsim"]


class CpuVec:
    backend = "oracle"

    def __init__(self, n, game_mode=0, rule_bits=A.RULE_DEFAULT_TENHOU, seeds=None, seed_base=0, log_cap_words=0, device=0):
        assert n == 1, "the CPU stand-in drives one game"
        self.n, self.game_mode = 1, int(game_mode)
        self.rule_bits, self.log_cap_words = int(rule_bits), int(log_cap_words)
        seed = int(seeds[0]) if seeds is not None else int(seed_base)
        self.seed = seed
        if self.backend == "oracle":
            import oracle

            self.lib = oracle.load()
            self.h = self.lib.orc_game_new(self.game_mode, seed, 0, self.rule_bits, 1)
            self._p = "orc_"
        else:
            self.lib = _hostsim().load()
            self.h = self.lib.hs_game_new(self.game_mode, seed, self.rule_bits, 1 << 16)
            self._p = "hs_"

    def _f(self, name):
        return getattr(self.lib, self._p + name)

    def reset(self, oya=None, round_wind=None, honba=None, kyotaku=None, scores=None, walls=None):
        wall = None if walls is None else list(walls[0])
        w = (C.c_uint8 * len(wall))(*wall) if wall is not None else None
        s = None
        if scores is not None:
            sc = list(scores[0])
            s = (C.c_int32 * 4)(*(sc + [0] * (4 - len(sc))))
        self._f("game_reset")(self.h, int(oya or 0), int(round_wind or 0), int(honba or 0), int(kyotaku or 0), w, s)

    def get_state(self, game=0):
        s = A.GameState()
        self._f("game_snapshot")(self.h, C.byref(s))
        return s

    def set_state(self, game, s):
        if self.backend == "hostsim":     # the product refuses records that are not positions (rv_vec_set_state)
            why = self.lib.hs_state_defect(C.byref(s))
            if why:
                raise ValueError(why.decode())
        self._f("game_load_snapshot")(self.h, C.byref(s))

    def step(self, arr):
        self._f("game_step")(self.h, arr)

    def legal_actions(self):
        acts = (A.Action * (A.NP * A.MAX_LEGAL))()
        counts = np.zeros((1, A.NP), np.uint8)
        for p in range(A.NP):
            sub = C.cast(C.byref(acts, C.sizeof(A.Action) * p * A.MAX_LEGAL), C.POINTER(A.Action))
            counts[0, p] = self._f("game_legal")(self.h, p, sub)
        return acts, counts

    def events(self, game=0):
        n = self._f("game_events")(self.h, None, 0)
        buf = (C.c_uint32 * max(n, 1))()
        self._f("game_events")(self.h, buf, n)
        return list(buf[:n])

    def mjai_log(self, game=0, viewer=-1, skip_events=0):
        """oracle: its own text log, written at event time (oracle/json.hpp); hostsim: the product's renderer over the words"""
        if self.backend == "oracle":
            n = self.lib.orc_game_mjai_log(self.h, viewer, None, 0)
            buf = C.create_string_buffer(n + 1)
            self.lib.orc_game_mjai_log(self.h, viewer, buf, n + 1)
            return (buf.value.decode().split("\n") if n else [])[skip_events:]
        from riichienv_b200._lib import events_to_json

        words, i = self.events(), 0
        for _ in range(skip_events):
            if i >= len(words):
                break
            i += max(1, (int(words[i]) >> 8) & 0xFF)
        return events_to_json(words[i:], viewer)

    def apply_events(self, events):
        self._f("game_apply_event")(self.h, events)

    def replay_begin(self, kyokus):
        self._f("game_replay_begin")(self.h, kyokus)

    def apply_log_actions(self, actions):
        self._f("game_apply_log_action")(self.h, actions)

    def call(self, op):
        """env.rs:624-631 hooks: op 0 reveal_kan_dora -> indicator count; op 1 -> list of ura indicator tile ids"""
        out = (C.c_uint8 * 8)()
        n = self._f("game_call")(self.h, int(op), out)
        return list(out[:n]) if op == 1 else n

    def clone(self):
        o = type(self)(1, self.game_mode, self.rule_bits, seeds=[self.seed], log_cap_words=self.log_cap_words)
        o.set_state(0, self.get_state(0))
        self._f("game_copy_log")(o.h, self.h)
        return o

    # ---- tensors (oracle only; the GPU suite covers the kernels) ----
    def _owes(self, pid):
        _, counts = self.legal_actions()
        if not counts[0, pid]:
            raise ValueError(f"seat {pid} owes no action; the tensors are defined for the observations step()/reset() return")

    def encode_single(self, pid, extended=False):
        self._owes(pid)
        w = 27 if self.game_mode >= 3 else 34
        a = np.zeros((215 if extended else 74, w), np.float32)
        if extended:
            self._f("game_encode_ext")(self.h, pid, a.ctypes.data_as(C.POINTER(C.c_float)))
        else:
            m = np.zeros(82, np.uint8)
            self._f("game_encode")(self.h, pid, a.ctypes.data_as(C.POINTER(C.c_float)), m.ctypes.data_as(C.POINTER(C.c_uint8)))
        return a.tobytes()

    def encode_kawa_single(self, pid):
        self._owes(pid)
        a = np.zeros((3, 7, 27) if self.game_mode >= 3 else (4, 7, 34), np.float32)
        self._f("game_encode_kawa")(self.h, a.ctypes.data_as(C.POINTER(C.c_float)))
        return a.tobytes()

    def encode_seq_single(self, pid, start_word):
        self._owes(pid)
        sp, nu = np.zeros(25, np.uint16), np.zeros(12, np.float32)
        pr, ca, le = np.zeros((512, 5), np.uint16), np.zeros((64, 4), np.uint16), np.zeros(3, np.uint16)
        p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
        fn = self.lib.orc_game_encode_seq if self.backend == "oracle" else self.lib.hs_game_encode_seq
        fn(self.h, pid, int(start_word), len(self.events()), 1, p16(sp),
                                     nu.ctypes.data_as(C.POINTER(C.c_float)), p16(pr), 512, p16(ca), p16(le))
        return sp, nu, pr, ca, le


class OracleVec(CpuVec):
    backend = "oracle"


class HostsimVec(CpuVec):
    backend = "hostsim"


def eval_queries_cpu(backend):
    """stand-in for riichienv_b200.hand.eval_queries (rv_hand_eval_batch) on the CPU checkers"""
    def run(queries, device=0):
        n = len(queries)
        arr = (A.HandQuery * n)(*queries)
        out = (A.HandResult * n)()
        if backend == "oracle":
            import oracle

            oracle.load().orc_hand_eval(arr, out, n)
        else:
            _hostsim().load().hs_hand_eval(arr, out, n)
        return out
    return run
