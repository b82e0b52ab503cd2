"""Attribute-dict inputs of the model (reference: utils/instance.py:9-171).

The model reads ``items.region_features`` / ``grid_features`` / ``region_boxes`` /
``caption_tokens``; ``InstanceList`` zero-pads variable-length per-image tensors into a batch.
"""

from __future__ import annotations

from collections import OrderedDict
from typing import Any, List

import numpy as np
import torch


class Instance(OrderedDict):
    def __init__(self, **kwargs):
        super().__init__(kwargs)

    def __setattr__(self, key, value):
        self[key] = value

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)

    def get_fields(self) -> List[str]:
        return list(self.keys())


def _pad_to_longest(values: List[torch.Tensor]) -> List[torch.Tensor]:
    """Zero-pad dim 0 of every tensor to the longest one and add a leading batch dim."""
    longest = max(v.shape[0] for v in values) if values[0].dim() > 0 else 0
    padded = []
    for v in values:
        if v.dim() > 0 and v.shape[0] < longest:
            pad = torch.zeros((longest - v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype)
            v = torch.cat([v, pad], dim=0)
        padded.append(v.unsqueeze(0))
    return padded


class InstanceList(OrderedDict):
    def __init__(self, instance_list: List[Instance] = ()):
        super().__init__()
        if len(instance_list) == 0:
            return
        assert all(isinstance(i, Instance) for i in instance_list)
        for key in instance_list[0].get_fields():
            values = [inst.get(key) for inst in instance_list]
            if isinstance(values[0], np.ndarray):
                values = [torch.tensor(v) for v in values]
            if isinstance(values[0], torch.Tensor):
                values = torch.cat(_pad_to_longest(values), dim=0)
            self[key] = values

    def __setattr__(self, name: str, val: Any) -> None:
        if name.startswith("_"):
            super().__setattr__(name, val)
        else:
            self[name] = val

    def __getattr__(self, name: str) -> Any:
        if name.startswith("_") or name not in self:
            return None
        return self[name]

    def set(self, name: str, value: Any) -> None:
        self[name] = value

    def has(self, name: str) -> bool:
        return name in self

    def get_fields(self) -> List[str]:
        return list(self.keys())

    @property
    def batch_size(self) -> int:
        for value in self.values():
            if isinstance(value, torch.Tensor):
                return value.shape[0]
            if isinstance(value, list):
                return len(value)
        return 0

    def to(self, *args: Any, **kwargs: Any) -> "InstanceList":
        ret = InstanceList()
        for key, value in self.items():
            ret[key] = value.to(*args, **kwargs) if hasattr(value, "to") else value
        return ret
