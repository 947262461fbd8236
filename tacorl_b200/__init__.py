"""tacorl_b200 — B200-native (sm_100a) training hot path for PlayLMP and TACO-RL.

Drop-in mirror of the reference's Hydra-configured network / module API
(`tacorl.networks.*`, `tacorl.modules.*` -> `tacorl_b200.networks.*`, `tacorl_b200.modules.*`),
with every tensor op executed by hand-written CUDA behind the C ABI in include/tacorl_b200.h.
"""
__version__ = "0.1.0"
