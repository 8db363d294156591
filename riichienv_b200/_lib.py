"""ctypes binding of libriichienv_b200.so (the C ABI in include/riichienv_b200.h).

The product has no CPU fallback: if the CUDA library is missing this module raises
ImportError loudly, and every compute call fails with RuntimeError when no GPU is
visible (RV_ERR_CUDA).
"""
import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RV_LIB_PATH") or os.path.join(_HERE, "libriichienv_b200.so")  # override: A/B builds
_LIB = None


class RvError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). riichienv_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp = C.c_void_p
    L.rv_last_error.restype = C.c_char_p
    L.rv_ctx_create.argtypes = [C.c_int, P(vp)]
    L.rv_ctx_destroy.argtypes = [vp]
    L.rv_ctx_sync.argtypes = [vp]
    L.rv_ctx_stream.restype = vp
    L.rv_ctx_stream.argtypes = [vp]
    L.rv_timer_mark.argtypes = [vp, C.c_int]
    L.rv_timer_elapsed.argtypes = [vp, C.c_int, C.c_int, P(C.c_float)]
    L.rv_vec_reseed.argtypes = [vp, P(C.c_uint64), C.c_uint64]
    L.rv_hand_eval_batch.argtypes = [vp, P(A.HandQuery), P(A.HandResult), C.c_int64]
    L.rv_hand_eval_batch_device.argtypes = [vp, vp, vp, C.c_int64]
    L.rv_hand_queries_seeded.argtypes = [vp, vp, C.c_uint64, C.c_int64]
    L.rv_hand_query_seeded_host.argtypes = [C.c_uint64, P(A.HandQuery)]
    L.rv_calculate_score.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, P(C.c_uint32)]
    L.rv_wall_from_seed.argtypes = [C.c_uint64, C.c_uint64, C.c_int, P(C.c_uint8)]
    L.rv_vec_create.argtypes = [vp, C.c_int64, C.c_int, C.c_uint32, P(C.c_uint64), C.c_uint64, C.c_uint32, P(vp)]
    L.rv_vec_destroy.argtypes = [vp]
    L.rv_vec_size.restype = C.c_int64
    L.rv_vec_size.argtypes = [vp]
    L.rv_vec_reset.argtypes = [vp, P(C.c_uint8), P(C.c_uint8), P(C.c_uint8), P(C.c_uint32), P(C.c_int32), P(C.c_uint8)]
    L.rv_vec_legal_actions.argtypes = [vp, P(A.Action), P(C.c_uint8)]
    L.rv_vec_step.argtypes = [vp, P(A.Action)]
    L.rv_vec_step_random.argtypes = [vp, C.c_uint64, C.c_uint32, P(C.c_uint64)]
    L.rv_vec_step_agent.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint32, P(C.c_uint64)]
    L.rv_vec_step_random_async.argtypes = [vp, C.c_uint64, C.c_uint32]
    L.rv_vec_steps_total.argtypes = [vp, P(C.c_uint64), P(C.c_int64)]
    L.rv_vec_results.argtypes = [vp, P(C.c_uint8), P(C.c_int32), P(C.c_uint8)]
    L.rv_vec_counters.argtypes = [vp, P(C.c_uint32), P(C.c_uint32), P(C.c_uint32), P(C.c_uint64)]
    L.rv_vec_get_state.argtypes = [vp, C.c_int64, P(A.GameState)]
    L.rv_vec_set_state.argtypes = [vp, C.c_int64, P(A.GameState)]
    L.rv_vec_state_device_ptr.argtypes = [vp, P(vp)]
    L.rv_vec_clone.argtypes = [vp, P(vp)]
    L.rv_vec_debug_call.argtypes = [vp, C.c_int64, C.c_int, P(C.c_uint8), P(C.c_int)]
    L.rv_vec_events.argtypes = [vp, C.c_int64, P(C.c_uint32), C.c_uint32, P(C.c_uint32)]
    L.rv_event_to_json.argtypes = [P(C.c_uint32), C.c_uint32, C.c_int, C.c_char_p, C.c_uint32]
    L.rv_vec_encode.argtypes = [vp, vp, vp, vp, C.c_int64, P(C.c_int64)]
    L.rv_vec_encode_ext.argtypes = [vp, vp, vp, vp, C.c_int64, P(C.c_int64)]
    L.rv_vec_encode_kawa.argtypes = [vp, vp, vp, C.c_int64, P(C.c_int64)]
    L.rv_vec_observe_step_random.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_int64, P(C.c_int64)]
    L.rv_vec_encode_seq.argtypes = [vp, C.c_int, P(C.c_uint32), vp, vp, vp, C.c_int, vp, vp, vp, C.c_int64, P(C.c_int64)]
    L.rv_multi_create.argtypes = [P(C.c_int), C.c_int, C.c_int64, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, P(vp)]
    L.rv_multi_destroy.argtypes = [vp]
    L.rv_multi_devices.argtypes = [vp]
    L.rv_multi_size.restype = C.c_int64
    L.rv_multi_size.argtypes = [vp]
    L.rv_multi_shard.argtypes = [vp, C.c_int, P(vp), P(C.c_int64), P(C.c_int64)]
    L.rv_multi_reset.argtypes = [vp]
    L.rv_multi_reseed.argtypes = [vp, C.c_uint64]
    L.rv_multi_step_random.argtypes = [vp, C.c_uint64, C.c_uint32, P(C.c_uint64)]
    L.rv_multi_results.argtypes = [vp, P(C.c_uint8), P(C.c_int32), P(C.c_uint8)]
    L.rv_multi_counters.argtypes = [vp, P(C.c_uint32), P(C.c_uint32), P(C.c_uint32), P(C.c_uint64)]
    L.rv_multi_stats.argtypes = [vp, P(A.RunStats)]
    L.rv_sizeof.argtypes = [C.c_int]
    L.rv_vec_apply_events.argtypes = [vp, P(A.MjaiEvent)]
    L.rv_replay_from_jsonl.argtypes = [C.c_char_p, C.c_uint32, P(vp)]
    L.rv_replay_from_text.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, P(vp)]
    L.rv_replay_from_mjsoul_json.argtypes = [C.c_char_p, C.c_uint32, P(vp)]
    L.rv_replay_from_mjsoul_text.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, P(vp)]
    L.rv_replay_free.argtypes = [vp]
    L.rv_replay_num_rounds.argtypes = [vp]
    L.rv_replay_kyoku.argtypes = [vp, C.c_int, P(A.LogKyoku)]
    L.rv_replay_actions.argtypes = [vp, C.c_int, P(A.LogAction), C.c_int, P(C.c_int)]
    L.rv_replay_paishan.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, P(C.c_int)]
    L.rv_replay_win_contexts.argtypes = [vp, C.c_int, P(A.WinContext), C.c_int, P(C.c_int)]
    L.rv_replay_from_files.argtypes = [P(C.c_char_p), C.c_int, C.c_int, C.c_uint32, C.c_int, P(vp), P(C.c_int)]
    L.rv_replay_totals.argtypes = [vp, C.c_int, P(C.c_int64), P(C.c_int64)]
    L.rv_replay_flatten.argtypes = [vp, C.c_int, P(A.LogKyoku), P(A.LogAction), P(C.c_int64), P(C.c_int32)]
    L.rv_replay_own_turn_labels.argtypes = [P(A.LogAction), C.c_int64, C.c_int, P(C.c_int16), P(C.c_int16)]
    L.rv_replay_progression.argtypes = [P(A.LogAction), P(C.c_uint8), C.c_int, P(C.c_uint16), C.c_int, P(C.c_int)]
    L.rv_replay_actions_aux.argtypes = [vp, C.c_int, P(A.LogActionAux), C.c_int, P(C.c_int)]
    if L.rv_replay_sizeof(0) != C.sizeof(A.WinContext) or L.rv_replay_sizeof(1) != C.sizeof(A.LogActionAux):
        raise RuntimeError("ABI mismatch: rv_win_context / rv_log_action_aux")
    L.rv_vec_replay_begin.argtypes = [vp, P(A.LogKyoku)]
    L.rv_vec_apply_log_actions.argtypes = [vp, P(A.LogAction)]
    L.rv_vec_replay_load.argtypes = [vp, P(A.LogKyoku), P(A.LogAction), P(C.c_int64)]
    L.rv_vec_replay_advance.argtypes = [vp, P(C.c_int64)]
    for i, T in enumerate((A.GameState, A.HandQuery, A.HandResult, A.Action, A.MjaiEvent, A.RunStats, A.LogAction, A.LogKyoku)):
        if L.rv_sizeof(i) != C.sizeof(T):
            raise ImportError(f"ABI mismatch for {T.__name__}: C {L.rv_sizeof(i)} != ctypes {C.sizeof(T)}")
    _LIB = L
    return L


def check(rc: int):
    if rc == 0:
        return
    msg = lib().rv_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    raise RvError(f"riichienv_b200 error {rc}: {msg}")


class Context:
    """One CUDA device + stream + lookup tables (rv_ctx)."""

    _cache = {}

    def __init__(self, device: int = 0):
        self.handle = C.c_void_p()
        check(lib().rv_ctx_create(int(device), C.byref(self.handle)))
        self.device = int(device)

    @classmethod
    def get(cls, device: int = 0) -> "Context":
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    def sync(self):
        check(lib().rv_ctx_sync(self.handle))

    def timer_mark(self, idx: int):
        check(lib().rv_timer_mark(self.handle, idx))

    def timer_elapsed(self, a: int, b: int) -> float:
        ms = C.c_float(0)
        check(lib().rv_timer_elapsed(self.handle, a, b, C.byref(ms)))
        return float(ms.value)

    @property
    def stream(self) -> int:
        return int(lib().rv_ctx_stream(self.handle) or 0)


def events_to_json(words, viewer: int = -1):
    """Render a binary event stream (sequence of u32) as a list of MJAI JSON strings."""
    L = lib()
    n = len(words)
    arr = (C.c_uint32 * n)(*words) if not isinstance(words, C.Array) else words
    buf = C.create_string_buffer(2048)
    out = []
    i = 0
    while i < n:
        sub = C.cast(C.byref(arr, 4 * i), C.POINTER(C.c_uint32))
        used = L.rv_event_to_json(sub, n - i, viewer, buf, 2048)
        if used <= 0:
            raise ValueError(f"malformed event stream at word {i}")
        out.append(buf.value.decode())
        i += used
    return out
