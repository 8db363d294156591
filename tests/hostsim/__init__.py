"""TEST INFRASTRUCTURE ONLY — host compile of the CUDA device sources (see cuda_shim.h).

Used by CPU-side tests to diff the kernel logic against the oracle where no GPU exists.
Never imported by the riichienv_b200 package.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    from riichienv_b200 import _abi as A

    so = os.path.join(_HERE, "libhostsim.so")
    csrc = os.path.join(_HERE, "..", "..", "riichienv_b200", "csrc")
    srcs = [os.path.join(_HERE, "hostsim.cpp"), os.path.join(_HERE, "cuda_shim.h")] + [
        os.path.join(csrc, f) for f in ("game.cuh", "hand.cuh", "tables.cuh", "obs.cuh", "obs_ext.cuh", "obs_ext3.cuh", "seq.cuh", "validate.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", so,
                               os.path.join(_HERE, "hostsim.cpp")])
    lib = C.CDLL(so)
    P = C.POINTER
    lib.hs_init.argtypes = [C.c_char_p, C.c_int]
    lib.hs_hand_eval.argtypes = [P(A.HandQuery), P(A.HandResult), C.c_int64]
    lib.hs_shanten_counts.argtypes = [P(C.c_uint8), C.c_int]
    lib.hs_shanten_counts_3p.argtypes = [P(C.c_uint8), C.c_int]
    lib.hs_is_agari.argtypes = [P(C.c_uint8)]
    lib.hs_waits.restype = C.c_uint64
    lib.hs_waits.argtypes = [P(C.c_uint8)]
    lib.hs_game_new.restype = C.c_void_p
    lib.hs_game_new.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
    lib.hs_game_free.argtypes = [C.c_void_p]
    lib.hs_game_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, P(C.c_uint8), P(C.c_int32)]
    lib.hs_game_legal.argtypes = [C.c_void_p, C.c_int, P(A.Action)]
    lib.hs_game_step.argtypes = [C.c_void_p, P(A.Action)]
    lib.hs_game_random_step.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    lib.hs_game_random_step_coopdeal.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    lib.hs_game_apply_event.argtypes = [C.c_void_p, P(A.MjaiEvent)]
    lib.hs_game_agent_step.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
    lib.hs_game_random_step_deferred.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    lib.hs_game_random_step_deferred.restype = C.c_int
    lib.hs_game_snapshot.argtypes = [C.c_void_p, P(A.GameState)]
    lib.hs_game_apply_log_action.argtypes = [C.c_void_p, P(A.LogAction)]
    lib.hs_game_replay_begin.argtypes = [C.c_void_p, P(A.LogKyoku)]
    lib.hs_game_load_snapshot.argtypes = [C.c_void_p, P(A.GameState)]
    lib.hs_state_defect.restype = C.c_char_p
    lib.hs_state_defect.argtypes = [P(A.GameState)]
    lib.hs_game_call.argtypes = [C.c_void_p, C.c_int, P(C.c_uint8)]
    lib.hs_game_copy_log.argtypes = [C.c_void_p, C.c_void_p]
    lib.hs_game_events.restype = C.c_uint32
    lib.hs_game_events.argtypes = [C.c_void_p, P(C.c_uint32), C.c_uint32]
    lib.hs_game_encode.argtypes = [C.c_void_p, C.c_int, P(C.c_float), P(C.c_uint8)]
    lib.hs_game_encode_ext.argtypes = [C.c_void_p, C.c_int, P(C.c_float)]
    lib.hs_game_encode_kawa.argtypes = [C.c_void_p, P(C.c_float)]
    lib.hs_game_encode_seq.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, P(C.c_uint16), P(C.c_float),
                                       P(C.c_uint16), C.c_int, P(C.c_uint16), P(C.c_uint16)]
    lib.hs_wall_from_seed.argtypes = [C.c_uint64, C.c_uint64, C.c_int, P(C.c_uint8)]
    # table cache beside the library (git-ignored), named after the sources that define the tables and checksummed
    import hashlib

    tag = hashlib.sha1(b"".join(open(os.path.join(csrc, f), "rb").read() for f in ("tables.cuh", "hand.cuh"))).hexdigest()[:12]
    cache = os.environ.get("RV_HOSTSIM_CACHE", os.path.join(_HERE, f"_tables_{tag}.bin"))
    for old in os.listdir(_HERE):
        if old.startswith("_tables_") and old.endswith(".bin") and old != os.path.basename(cache):
            os.remove(os.path.join(_HERE, old))
    lib.hs_init(cache.encode(), os.cpu_count() or 1)
    _LIB = lib
    return lib
