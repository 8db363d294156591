from riichienv_b200.env import Action, Action3P, ActionType  # noqa: F401
