"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for oracle/liboracle.so (the CPU restatement of the reference).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product
(riichienv_b200/) never does.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "game.hpp", "hand.hpp", "wall.hpp", "shanten.hpp", "obs.hpp", "seq.hpp", "json.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "riichienv_b200.h"))
    stale = not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    from riichienv_b200 import _abi as A

    lib = C.CDLL(build())
    P = C.POINTER
    lib.orc_hand_eval.argtypes = [P(A.HandQuery), P(A.HandResult), C.c_int64]
    lib.orc_hand_eval_mt.argtypes = [P(A.HandQuery), P(A.HandResult), C.c_int64, C.c_int]
    lib.orc_hand_queries_seeded.argtypes = [P(A.HandQuery), C.c_uint64, C.c_int64]
    lib.orc_is_agari.argtypes = [P(C.c_uint8)]
    lib.orc_is_tenpai_counts.argtypes = [P(C.c_uint8)]
    lib.orc_shanten_counts.argtypes = [P(C.c_uint8), C.c_int]
    lib.orc_shanten_counts_3p.argtypes = [P(C.c_uint8), C.c_int]
    lib.orc_calculate_score.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, P(C.c_uint32)]
    lib.orc_wall_from_seed.argtypes = [C.c_uint64, C.c_uint64, C.c_int, P(C.c_uint8)]
    lib.orc_chacha_words.argtypes = [P(C.c_uint32), C.c_int, C.c_int, P(C.c_uint32)]
    lib.orc_run_agent_walls.restype = C.c_int64
    lib.orc_run_agent_walls.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, P(C.c_uint8),
                                        P(C.c_int32), P(C.c_uint8), P(C.c_uint8), P(C.c_uint32), P(C.c_uint32), P(C.c_uint32), P(C.c_uint64)]
    lib.orc_game_new.restype = C.c_void_p
    lib.orc_game_new.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_uint32, C.c_int]
    lib.orc_game_free.argtypes = [C.c_void_p]
    lib.orc_game_reset.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, P(C.c_uint8), P(C.c_int32)]
    lib.orc_game_legal.argtypes = [C.c_void_p, C.c_int, P(A.Action)]
    lib.orc_game_step.argtypes = [C.c_void_p, P(A.Action)]
    lib.orc_game_random_step.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
    lib.orc_games_legal_batch.argtypes = [P(C.c_void_p), C.c_int64, P(A.Action), P(C.c_uint8)]
    lib.orc_games_random_step_batch.argtypes = [P(C.c_void_p), C.c_int64, C.c_uint64, C.c_uint64]
    lib.orc_game_snapshot.argtypes = [C.c_void_p, P(A.GameState)]
    lib.orc_game_apply_log_action.argtypes = [C.c_void_p, P(A.LogAction)]
    lib.orc_game_replay_begin.argtypes = [C.c_void_p, P(A.LogKyoku)]
    lib.orc_game_load_snapshot.argtypes = [C.c_void_p, P(A.GameState)]
    lib.orc_game_call.argtypes = [C.c_void_p, C.c_int, P(C.c_uint8)]
    lib.orc_game_copy_log.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_game_mjai_log.restype = C.c_uint32
    lib.orc_game_mjai_log.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_uint32]
    lib.orc_game_events.restype = C.c_uint32
    lib.orc_game_events.argtypes = [C.c_void_p, P(C.c_uint32), C.c_uint32]
    lib.orc_game_encode.argtypes = [C.c_void_p, C.c_int, P(C.c_float), P(C.c_uint8)]
    lib.orc_games_encode_batch.restype = C.c_int64
    lib.orc_games_encode_batch.argtypes = [P(C.c_void_p), C.c_int64, P(C.c_float), P(C.c_uint8), P(C.c_int32), C.c_int64]
    lib.orc_game_encode_ext.argtypes = [C.c_void_p, C.c_int, P(C.c_float)]
    lib.orc_game_encode_kawa.argtypes = [C.c_void_p, P(C.c_float)]
    lib.orc_ukeire.argtypes = [P(C.c_int), C.c_int, P(C.c_int), C.c_int, P(C.c_int)]
    lib.orc_ukeire_3p.argtypes = [P(C.c_int), C.c_int, P(C.c_int), C.c_int, P(C.c_int)]
    lib.orc_game_encode_seq.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_int, P(C.c_uint16), P(C.c_float),
                                        P(C.c_uint16), C.c_int, P(C.c_uint16), P(C.c_uint16)]
    for name in ("orc_seq_encode_chi", "orc_seq_encode_pon"):
        getattr(lib, name).argtypes = [C.c_int, C.c_int, C.c_int]
    lib.orc_seq_kan37.argtypes = [C.c_int]
    lib.orc_seq_relative_from.argtypes = [C.c_int, C.c_int]
    lib.orc_run_random.restype = C.c_int64
    lib.orc_run_random.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint32, C.c_int,
                                   P(C.c_int32), P(C.c_uint8), P(C.c_uint8), P(C.c_uint32), P(C.c_uint32),
                                   P(C.c_uint32), P(C.c_uint64), P(C.c_uint32)]
    lib.orc_run_agent.restype = C.c_int64
    lib.orc_run_agent.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.c_int64, C.c_uint64, C.c_uint32, C.c_int,
                                  P(C.c_int32), P(C.c_uint8), P(C.c_uint8), P(C.c_uint32), P(C.c_uint32),
                                  P(C.c_uint32), P(C.c_uint64), P(C.c_uint64)]
    lib.orc_game_apply_event.argtypes = [C.c_void_p, P(A.MjaiEvent)]
    lib.orc_game_agent_step.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64]
    _LIB = lib
    return lib
