"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (ErickRosete/tacorl) from
/root/reference/src through the dependency stubs in oracle/stubs.

Only usable in the build container (the reference does not travel to the GPU box).
Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py to pin the
restatement in oracle/tacorl_oracle.py against the real reference code.

Nothing in the product package (tacorl_b200/) imports this file.
"""
import copy
import os
import sys
import types

REF_SRC = "/root/reference/src"
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def reference_available():
    return os.path.isdir(os.path.join(REF_SRC, "tacorl"))


def import_reference():
    """Put stubs + reference on sys.path and return the `tacorl` package."""
    if not reference_available():
        raise RuntimeError("reference not present at /root/reference")
    for p in (REF_SRC, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    # plan_recognition_net.py:11 imports a module that does not exist in the reference.
    if "tacorl.utils.test" not in sys.modules:
        import tacorl.utils  # noqa: F401

        m = types.ModuleType("tacorl.utils.test")
        m.get_simulated_trajectory = lambda *a, **k: None
        sys.modules["tacorl.utils.test"] = m
    import tacorl

    return tacorl


from .ref_loader_cfg import *  # noqa: F401,F403,E402
from .ref_loader_cfg import play_lmp_cfg, tacorl_cfg  # noqa: E402


def build_reference_play_lmp(**kw):
    import_reference()
    from tacorl.modules.play_lmp.play_lmp_for_rl import PlayLMP

    return PlayLMP(**copy.deepcopy(play_lmp_cfg(**kw)))


def build_reference_tacorl(play_lmp, **overrides):
    """TACORL around an in-memory PlayLMP (checkpoint loader monkey-patched)."""
    import_reference()
    import tacorl.modules.tacorl.tacorl as T

    T.load_pl_module_from_checkpoint = lambda *a, **k: play_lmp
    cfg = tacorl_cfg()
    cfg.update(overrides)
    return T.TACORL(**copy.deepcopy(cfg))
