"""Test-infrastructure stub of `hydra` (not installed in this image).

Only what the reference's hot-path modules touch at import/instantiate time.
Used solely by oracle/ref_loader.py to import /root/reference/src unmodified.
"""
from . import utils  # noqa: F401


def main(*args, **kwargs):
    def deco(fn):
        return fn

    return deco
