from riichienv import *  # noqa: F401,F403
from riichienv import Action, Observation, Observation3P, RiichiEnv  # noqa: F401
