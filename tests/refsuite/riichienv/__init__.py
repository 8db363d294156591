"""TEST INFRASTRUCTURE ONLY: the name `riichienv` for the reference's own pytest suite, bound to riichienv_b200.

RV_REFSUITE_BACKEND selects what executes the game logic behind the shim:
  gpu (default when a CUDA device is visible)  the product, libriichienv_b200.so through its C ABI
  oracle / hostsim                             CPU checkers (tests/refsuite/cpu_vec.py) for the GPU-less authoring box
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import riichienv_b200 as _rb  # noqa: E402
from riichienv_b200 import convert, env as _env, hand as _hand  # noqa: E402


def _pick_backend():
    b = os.environ.get("RV_REFSUITE_BACKEND")
    if b:
        return b
    try:
        import torch

        return "gpu" if torch.cuda.is_available() else "oracle"
    except Exception:
        return "oracle"


BACKEND = _pick_backend()
if BACKEND != "gpu":
    import importlib.util as _ilu

    _spec = _ilu.spec_from_file_location("rv_cpu_vec", os.path.join(_ROOT, "tests", "refsuite", "cpu_vec.py"))
    _cv = _ilu.module_from_spec(_spec)
    sys.modules["rv_cpu_vec"] = _cv
    _spec.loader.exec_module(_cv)

    _env.VecRiichiEnv = {"oracle": _cv.OracleVec, "hostsim": _cv.HostsimVec}[BACKEND]
    _hand.eval_queries = _cv.eval_queries_cpu(BACKEND)

from riichienv_b200.env import (Action, Action3P, ActionType, GameRule, GameType, Meld, MeldType, Observation,  # noqa: E402,F401
                                Observation3P, Phase, RiichiEnv, Wind)
from riichienv_b200.hand import (Conditions, HandEvaluator, HandEvaluator3P, Score, WinResult, calculate_score,  # noqa: E402,F401
                                 calculate_shanten, calculate_shanten_3p, check_riichi_candidates)
from riichienv_b200.yaku_table import Yaku, get_all_yaku, get_yaku_by_id  # noqa: E402,F401
from riichienv_b200.convert import parse_hand, parse_tile  # noqa: E402,F401

EAST, SOUTH, WEST, NORTH = Wind.East, Wind.South, Wind.West, Wind.North


from riichienv_b200.replay import (Kyoku, KyokuIterator, KyokuStepIterator, MjaiReplay, MjSoulReplay, WinResultContext,  # noqa: E402,F401  (replay ingestion)
                                   WinResultContextIterator)
from . import consts  # noqa: E402,F401


def __getattr__(name):
    # viewer, legacy pure-Python rule classes: out of scope; the tests that import them are
    # recorded as out of scope in tests/refsuite/expected.txt
    raise AttributeError(f"riichienv_b200 does not provide {name} (out of scope, SURVEY.md §2)")
