"""YAML -> attribute-dict config, loading the reference's config files unchanged.

Mirrors configs/utils.py:4-5 of the reference (``get_config(yaml_file) -> CfgNode``); yacs is not a
dependency here, ``CfgNode`` is a small recursive attribute dict with the subset of the yacs API the
path uses (attribute access, ``clone``, ``merge_from_dict``).
"""

from __future__ import annotations

import copy
from pathlib import Path

import yaml

CONFIG_DIR = Path(__file__).resolve().parent


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        for key, value in (init_dict or {}).items():
            self[key] = CfgNode(value) if isinstance(value, dict) else value

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def __setattr__(self, key, value):
        self[key] = CfgNode(value) if isinstance(value, dict) and not isinstance(value, CfgNode) else value

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def merge_from_dict(self, other: dict) -> "CfgNode":
        for key, value in other.items():
            if isinstance(value, dict) and isinstance(self.get(key), CfgNode):
                self[key].merge_from_dict(value)
            else:
                self[key] = CfgNode(value) if isinstance(value, dict) else value
        return self


def get_config(yaml_file) -> CfgNode:
    """Load a config; bare names resolve inside this package's ``configs/`` directory."""
    path = Path(yaml_file)
    if not path.exists() and (CONFIG_DIR / path.name).exists():
        path = CONFIG_DIR / path.name
    with open(path, "r") as handle:
        return CfgNode(yaml.load(handle, Loader=yaml.FullLoader))
