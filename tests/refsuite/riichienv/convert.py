from riichienv_b200.convert import *  # noqa: F401,F403
