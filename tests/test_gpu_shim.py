"""The drop-in Python surface (riichienv_b200.RiichiEnv & friends) on the GPU — reads like the reference's tests."""
import json

import pytest

pytestmark = pytest.mark.gpu


def test_readme_loop_and_logs():
    # README.md:50-62 of the reference, unchanged except for the import
    from riichienv_b200 import RiichiEnv
    from riichienv_b200.agents import RandomAgent

    agent = RandomAgent(seed=3)
    env = RiichiEnv(game_mode="4p-red-east", seed=11)
    obs_dict = env.reset()
    steps = 0
    while not env.done():
        actions = {pid: agent.act(obs) for pid, obs in obs_dict.items()}
        obs_dict = env.step(actions)
        steps += 1
        assert steps < 5000
    assert sorted(env.ranks()) == [1, 2, 3, 4]
    assert sum(env.scores()) + 1000 * env.riichi_sticks == 100000
    log = env.mjai_log
    assert log[0]["type"] == "start_game" and log[-1]["type"] == "end_game" and log[-2]["type"] == "end_kyoku"
    assert len(env.points("basic")) == 4
    with pytest.raises(ValueError):
        env.points("nope")


def test_initialization_like_reference():  # tests/env/test_riichienv.py:9-64
    from riichienv_b200 import Action, ActionType, Observation, Phase, RiichiEnv, tid_to_mjai

    env = RiichiEnv(seed=42)
    assert len(env.wall) > 0
    obs_dict = env.reset()
    assert len(env.wall) == 83
    assert [len(h) for h in env.hands] == [14, 13, 13, 13]
    assert env.melds[0] == [] and env.discards[0] == []
    assert env.current_player == 0 and env.turn_count == 0 and env.is_done is False and env.needs_tsumo is False
    assert list(obs_dict.keys()) == [0]
    o = obs_dict[0]
    assert isinstance(o, Observation) and o.player_id == 0 and len(o.hand) == 14
    assert [e["type"] for e in o.events] == ["start_game", "start_kyoku", "tsumo"]
    assert len(o.new_events()) == 3 and len(o.legal_actions()) == 14
    d = o.to_dict()
    assert d["legal_actions"][0]["type"] == 0 and d["legal_actions"][0]["consume_tiles"] == []
    assert o.select_action_from_mjai({"type": "dahai", "pai": tid_to_mjai(o.hand[0]), "actor": 0}) is not None
    # basic step (tests/env/test_riichienv.py:66-126)
    obs_dict = env.step({0: Action(ActionType.DISCARD, tile=o.hand[-1])})
    while env.phase == Phase.WaitResponse:
        obs_dict = env.step({pid: Action(ActionType.PASS) for pid in env.active_players})
    assert env.phase == Phase.WaitAct and env.current_player == 1 and list(obs_dict.keys()) == [1]
    o1 = obs_dict[1]
    assert o1.events[1]["tehais"][0][0] == "?" and o1.events[1]["tehais"][1][0] != "?" and o1.events[2]["pai"] == "?"
    assert [e["type"] for e in env.mjai_log][:5] == ["start_game", "start_kyoku", "tsumo", "dahai", "tsumo"]
    assert sum(o1.mask()) == len({a.encode() for a in o1.legal_actions()})


def test_setters_pon_priority():  # tests/env/rule_validation/test_claim_priority.py via the PyO3-style setters
    from riichienv_b200 import Action, ActionType, Phase, RiichiEnv

    env = RiichiEnv(seed=1, game_mode=0)
    env.reset()
    h = env.hands
    h[0] = sorted([57] + [2] * 12)
    h[1] = sorted([62, 65] + [0] * 11)
    h[2] = sorted([56, 58] + [1] * 11)
    h[3] = [12, 16, 19, 21, 48, 59, 64, 77, 81, 89, 104, 130, 133]
    env.hands = h
    env.active_players = [0]
    env.current_player = 0
    env.phase = Phase.WaitAct
    env.needs_tsumo = False
    env.drawn_tile = 100
    h = env.hands
    h[0] = sorted(h[0] + [100])
    env.hands = h
    env.step({0: Action(ActionType.DISCARD, tile=57)})
    assert env.phase == Phase.WaitResponse and env.active_players == [1, 2]
    env.step({1: Action(ActionType.CHI, tile=57, consume_tiles=[62, 65]), 2: Action(ActionType.PON, tile=57, consume_tiles=[56, 58])})
    assert env.phase == Phase.WaitAct and env.active_players == [2]
    assert env.mjai_log[-1]["type"] == "pon"


def test_illegal_action_returns_empty_dict():  # tests/env/test_illegal_actions.py:5-60
    from riichienv_b200 import Action, ActionType, RiichiEnv

    env = RiichiEnv(game_mode="4p-red-east", seed=42)
    env.reset()
    bad = 0
    while bad in env.hands[0]:
        bad += 1
    assert env.step({0: Action(ActionType.DISCARD, tile=bad)}) == {}
    assert not env.done()
    ry = [e for e in env.mjai_log if e["type"] == "ryukyoku"][-1]
    assert "Error: Illegal Action" in ry["reason"] and ry["deltas"] == [-12000, 4000, 4000, 4000]
    assert env.scores() == [13000, 29000, 29000, 29000]
    with pytest.raises(ValueError):
        env.reset(scores=[1, 2, 3])


def test_hand_front_end():  # tests/test_agari_calculator.py, tests/test_shanten.py, README.md:221-272
    from riichienv_b200 import Conditions, HandEvaluator, calculate_score, calculate_shanten
    from tests.helpers import parse_hand

    r = HandEvaluator(parse_hand("111m33p12s111666z")).calc(18 * 4 + 2 * 4, conditions=Conditions())
    assert r.is_win and (r.han, r.fu, r.ron_agari) == (5, 60, 12000) and sorted(r.yaku) == [8, 10, 11, 22]
    assert calculate_shanten(parse_hand("19m19p19s1234567z")) == 0
    assert calculate_shanten(parse_hand("111m111z222z333z44z")) == -1
    assert calculate_shanten(parse_hand("19m147p258s12345z")) == 5
    assert calculate_shanten(parse_hand("1111m111122233z")) == 1
    assert HandEvaluator(parse_hand("123m456p789s1122z")).get_waits() == [27, 28]
    assert calculate_score(3, 30, True, False).pay_ron == 5800


def test_sanma_shim():  # tests/env/test_sanma.py:50-118, 390-401, 480-500
    from riichienv_b200 import Action, ActionType, GameType, Phase, RiichiEnv
    from riichienv_b200.agents import RandomAgent

    env = RiichiEnv(game_mode=GameType.SAN_HANCHAN, seed=42)
    obs = env.reset()
    assert env.num_players == 3 and env.scores() == [35000, 35000, 35000]
    assert [len(h) for h in env.hands] == [14, 13, 13] and len(env.wall) == 68
    assert env.points("basic") == [40.0, 0.0, -40.0] and set(env.ranks()) == {1, 2, 3}
    agent = RandomAgent(seed=1)
    steps = 0
    while not env.done() and steps < 3000:
        for o in obs.values():
            assert all(a.action_type != ActionType.CHI for a in o.legal_actions())
        obs = env.step({pid: agent.act(o) for pid, o in obs.items()})
        steps += 1
    assert env.done()
    assert sum(env.scores()) + 1000 * env.riichi_sticks == 105000
    assert all(len(e["deltas"]) == 3 for e in env.mjai_log if e["type"] in ("hora", "ryukyoku"))
