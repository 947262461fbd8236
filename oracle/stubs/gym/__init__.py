"""Test-infrastructure stub of `gym` (import-time only)."""
from . import envs, spaces  # noqa: F401


class Env:
    pass


class Wrapper(Env):
    pass


def make(*a, **k):
    raise RuntimeError("gym stub")


def register(*a, **k):
    pass
