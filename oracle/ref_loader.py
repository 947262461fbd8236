"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (ErickRosete/tacorl) from
/root/reference/src through the dependency stubs in oracle/stubs.

Usable where the reference is present: /root/reference in the build container, or the copy
oracle/vendor_reference.py makes under oracle/_ref (git-ignored; it travels to the GPU box with
the gpurun snapshot).  Used by oracle/make_golden.py to pin the restatement in
oracle/tacorl_oracle.py against the real reference code, and by bench.py's reference arm.

Nothing in the product package (tacorl_b200/) imports this file.
"""
import copy
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")


def _ref_src():
    """The reference package root: /root/reference/src in the build container, else the tree vendored by
    oracle/vendor_reference.py (oracle/_ref, git-ignored, travels to the GPU box)."""
    roots = ("/root/reference/src", os.path.join(_HERE, "_ref"))
    if os.environ.get("TACORL_REF_VENDORED_ONLY"):      # tests: behave like the GPU box
        roots = roots[1:]
    for p in roots:
        if os.path.isdir(os.path.join(p, "tacorl")):
            return p
    return None


REF_SRC = _ref_src()


def reference_available():
    return _ref_src() is not None


def import_reference():
    """Put stubs + reference on sys.path and return the `tacorl` package."""
    if not reference_available():
        raise RuntimeError("reference not present (neither /root/reference nor oracle/_ref)")
    for p in (_ref_src(), _STUBS):
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


def build_reference_cql(**overrides):
    """The flat-CQL baseline, CQL_Offline as composed by config/experiment/cql_real_world.yaml."""
    import_reference()
    from tacorl.modules.cql.cql_offline_lightning import CQL_Offline
    from .ref_loader_cfg import cql_offline_cfg

    cfg = cql_offline_cfg()
    cfg.update(overrides)
    return CQL_Offline(**copy.deepcopy(cfg))
