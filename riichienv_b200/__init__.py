"""riichienv_b200 — B200-native batched Riichi mahjong simulator (hot path of smly/RiichiEnv).

Python is a thin shim: every game-logic call goes to libriichienv_b200.so (CUDA, sm_100a) through the
C ABI in include/riichienv_b200.h.  Importing this package does not need a GPU; computing does.
"""
from . import _abi  # noqa: F401

__all__ = ["RiichiEnv", "VecRiichiEnv", "MultiVecRiichiEnv", "Observation", "Observation3P", "Action", "Action3P", "ActionType", "Phase", "Meld", "MeldType", "GameRule",
           "GameType", "Wind", "HandEvaluator", "Conditions", "calculate_score", "calculate_shanten", "calculate_shanten_3p", "tid_to_mjai",
           "MjaiReplay", "MjSoulReplay", "Kyoku", "ReplayBatch", "WinResultContext", "WinResultContextIterator", "KyokuStepIterator", "KyokuIterator", "Yaku", "get_yaku_by_id", "get_all_yaku",
           "check_riichi_candidates"]


def __getattr__(name):  # lazy: keep `import riichienv_b200` light and GPU-free
    if name in ("RiichiEnv", "Observation", "Observation3P", "Action", "Action3P", "ActionType", "Phase", "Meld", "MeldType", "GameRule", "GameType", "Wind",
                "tid_to_mjai"):
        from . import env

        return getattr(env, name)
    if name in ("VecRiichiEnv", "MultiVecRiichiEnv"):
        from . import vec_env

        return getattr(vec_env, name)
    if name in ("Yaku", "get_yaku_by_id", "get_all_yaku"):
        from . import yaku_table

        return getattr(yaku_table, name)
    if name in ("HandEvaluator", "Conditions", "calculate_score", "calculate_shanten", "calculate_shanten_3p", "WinResult", "check_riichi_candidates"):
        from . import hand

        return getattr(hand, name)
    if name in ("MjaiReplay", "MjSoulReplay", "Kyoku", "ReplayBatch", "WinResultContext", "WinResultContextIterator", "KyokuStepIterator",
                "KyokuIterator"):
        from . import replay

        return getattr(replay, name)
    raise AttributeError(name)
