from .random_agent import KeyedRandomAgent, RandomAgent  # noqa: F401
