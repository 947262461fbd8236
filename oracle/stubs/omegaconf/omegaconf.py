from . import DictConfig, ListConfig, OmegaConf  # noqa: F401
