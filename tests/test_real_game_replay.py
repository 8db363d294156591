"""A real 12-kyoku game replayed through the environment — the reference-held answers of tests/data/126_204_0_mjai.jsonl.

tests/golden/real_game_126_204_0.json (tests/golden/make_real_game.py) holds, per kyoku, a wall reconstructed from the log
and the players' logged decisions.  The env is reset onto that wall (reset(wall=) -> load_wall) and fed the decisions through
the public API (Observation.select_action_from_mjai, RiichiEnv.step); every event it then logs must equal the log's line:
the draws and dora flips (the wall geometry: deal order, live wall, rinshan, indicator slots), the call / riichi sequencing,
and the settlement of 9 wins and 3 exhaustive draws — `deltas` (han / fu / yaku -> points, honba, riichi sticks, tenpai
payments) and `ura_markers`.  Runs on the oracle and on the kernel code (host compile); under -m gpu on the product."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = json.load(open(os.path.join(HERE, "golden", "real_game_126_204_0.json")))


def _env_module(backend):
    """the shim with the requested executor behind it (tests/refsuite/riichienv picks by RV_REFSUITE_BACKEND)"""
    import importlib
    import sys

    os.environ["RV_REFSUITE_BACKEND"] = backend
    sys.path.insert(0, os.path.join(HERE, "refsuite"))
    for name in [m for m in sys.modules if m == "riichienv" or m.startswith("riichienv.")]:
        del sys.modules[name]
    import riichienv_b200.env as E
    import riichienv_b200.hand as Hd
    from riichienv_b200.vec_env import VecRiichiEnv

    E.VecRiichiEnv = VecRiichiEnv          # undo a previous backend's patch
    importlib.reload(Hd)
    mod = importlib.import_module("riichienv")
    sys.path.pop(0)
    return mod


def _same(exp, got):
    """log line vs the env's event: the fixture's log omits keys the env always writes (hora.tsumo, ryukyoku.reason, ankan.pai)"""
    if exp["type"] != got["type"]:
        return False
    for k, v in exp.items():
        if k == "consumed":
            if sorted(v) != sorted(got.get(k, [])):
                return False
        elif got.get(k) != v:
            return False
    return True


def _play(rv, ky):
    env = rv.RiichiEnv(game_mode="4p-red-half", seed=1)
    obs = env.reset(oya=ky["oya"], wall=ky["wall"], round_wind=ky["bakaze"], scores=ky["scores"], honba=ky["honba"],
                    kyotaku=ky["kyotaku"])
    todo = list(ky["decisions"])
    guard = 0
    while todo and not env.done():
        guard += 1
        assert guard < 2000
        d = todo[0]
        ty, actor = d["type"], d["actor"]
        if env.phase == rv.Phase.WaitAct:
            cur = env.current_player
            assert cur == actor and ty in ("dahai", "reach", "ankan", "kakan", "hora"), (d, cur)
            a = obs[cur].select_action_from_mjai(d)
            assert a is not None, f"{d} not among {obs[cur].legal_actions()}"
            obs = env.step({cur: a})
            todo.pop(0)
        else:
            acts = {p: rv.Action(rv.ActionType.PASS) for p in env.active_players}
            if ty in ("pon", "chi", "daiminkan", "kan") or (ty == "hora" and d["target"] != actor):
                if actor in obs:
                    a = obs[actor].select_action_from_mjai(d)
                    if a is not None:
                        acts[actor] = a
                        todo.pop(0)
                        # a double ron: the next decision is another hora on the same discard
                        while todo and todo[0]["type"] == "hora" and ty == "hora" and todo[0]["actor"] in obs:
                            a2 = obs[todo[0]["actor"]].select_action_from_mjai(todo[0])
                            if a2 is None:
                                break
                            acts[todo[0]["actor"]] = a2
                            todo.pop(0)
            obs = env.step(acts)
    while not todo and env.phase == rv.Phase.WaitResponse and not _round_over(env):      # trailing pass (nobody claimed the last discard)
        obs = env.step({p: rv.Action(rv.ActionType.PASS) for p in env.active_players})
    return env


def _round_over(env):
    return any(e["type"] in ("hora", "ryukyoku") for e in env.mjai_log[2:])


def _check(rv):
    wins = draws = 0
    for k, ky in enumerate(DATA["kyokus"]):
        env = _play(rv, ky)
        log = env.mjai_log
        assert log[1]["type"] == "start_kyoku" and log[1]["dora_marker"] == ky["dora_marker"] and log[1]["scores"] == ky["scores"]
        got = [e for e in log[2:] if e["type"] not in ("end_kyoku", "end_game", "start_kyoku")]
        got = got[: len(ky["expected"])] if len(got) > len(ky["expected"]) else got     # (a next round the env dealt itself)
        # the env plays on after the round (hanchan mode deals the next round from its own seed): cut at this round's end
        end = next(i for i, e in enumerate(got) if e["type"] in ("hora", "ryukyoku"))
        tail_horas = [e for e in got[end:] if e["type"] == "hora"] if got[end]["type"] == "hora" else []
        got = got[:end] + (tail_horas if tail_horas else [got[end]])
        assert len(got) == len(ky["expected"]), f"kyoku {k}: {len(got)} events, log has {len(ky['expected'])}"
        for i, (e, g) in enumerate(zip(ky["expected"], got)):
            assert _same(e, g), f"kyoku {k} event {i}: log {e} vs env {g}"
        wins += sum(1 for e in got if e["type"] == "hora")
        draws += sum(1 for e in got if e["type"] == "ryukyoku")
    assert (wins, draws) == (9, 3)


@pytest.mark.parametrize("backend", ["oracle", "hostsim"])
def test_real_game_replay(backend):
    _check(_env_module(backend))


@pytest.mark.gpu
def test_real_game_replay_gpu():
    _check(_env_module("gpu"))
