from . import cm, pyplot  # noqa: F401
