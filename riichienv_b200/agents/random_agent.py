"""Agents.  RandomAgent mirrors src/riichienv/agents/random_agent.py:6-15 (CPython Mersenne Twister, call-order
dependent); KeyedRandomAgent is the counter-keyed agent the on-device rollout uses (SURVEY.md §8 d), reproducible
independent of dict order."""
import random

_M = (1 << 64) - 1


def mix64(z: int) -> int:
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return z ^ (z >> 31)


class RandomAgent:
    def __init__(self, seed=None):
        self._rng = random.Random(seed)

    def act(self, obs):
        return self._rng.choice(obs.legal_actions())


class KeyedRandomAgent:
    def __init__(self, agent_seed: int, game_id: int):
        self.agent_seed, self.game_id = agent_seed, game_id

    def act(self, obs, step_count: int):
        legal = obs.legal_actions()
        key = (self.agent_seed ^ ((self.game_id * 0x9E3779B97F4A7C15) & _M) ^ (step_count << 8) ^ obs.player_id) & _M
        return legal[mix64(key) % len(legal)]
