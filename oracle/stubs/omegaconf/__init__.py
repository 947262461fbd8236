"""Test-infrastructure stub of `omegaconf`: configs are plain dicts/lists."""
import copy

import yaml

DictConfig = dict
ListConfig = list


class OmegaConf:
    @staticmethod
    def to_container(cfg, resolve=True, **kw):
        return copy.deepcopy(cfg)

    @staticmethod
    def load(path):
        with open(path) as f:
            return yaml.safe_load(f)

    @staticmethod
    def create(obj=None):
        return copy.deepcopy(obj) if obj is not None else {}

    @staticmethod
    def to_yaml(cfg):
        return yaml.safe_dump(cfg)
