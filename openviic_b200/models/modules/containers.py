"""Stateful module container: the decode-state registry that beam search reorders.

API of the reference's models/modules/containers.py:5-78 -- ``register_state``, ``states``, ``apply_to_states``,
``enable_statefulness`` / ``disable_statefulness``, the ``statefulness`` context manager, ``_is_stateful`` -- on a
different mechanism: every container keeps one ordered table ``name -> pristine default``, and one generator walks
the container tree (a container, then its container children depth-first in registration order; plain
``nn.Module`` children and everything below them are skipped, as in the reference).  All operations are loops
over that walk.  States are registered buffers, so they appear in ``state_dict()`` exactly like the reference's.
"""

from __future__ import annotations

from collections import OrderedDict
from contextlib import contextmanager
from typing import Iterator, Optional

from torch import nn


class Module(nn.Module):
    def __init__(self):
        super().__init__()
        self._is_stateful = False
        self._decode_states = OrderedDict()   # state name -> default value (a detached copy, or None)

    def register_state(self, name: str, default) -> None:
        self._decode_states[name] = default if default is None else default.detach().clone()
        self.register_buffer(name, default)

    def _state_owners(self) -> Iterator["Module"]:
        yield self
        for child in self.children():
            if isinstance(child, Module):
                yield from child._state_owners()

    def states(self):
        for owner in self._state_owners():
            for name in owner._decode_states:
                yield owner._buffers[name]

    def apply_to_states(self, fn) -> None:
        """``state = fn(state)`` for every state of the tree, in the order ``states()`` yields them."""
        for owner in self._state_owners():
            for name in owner._decode_states:
                owner._buffers[name] = fn(owner._buffers[name])

    def _load_defaults(self, batch_size: Optional[int]) -> None:
        """Every state back to its default -- broadcast over a leading batch dimension when ``batch_size`` is given
        (decode mode), as registered otherwise."""
        for owner in self._state_owners():
            for name, default in owner._decode_states.items():
                value = None
                if default is not None:
                    value = default.detach().clone().to(owner._buffers[name].device)
                    if batch_size is not None:
                        value = value.unsqueeze(0).expand(batch_size, *value.shape).contiguous()
                owner._buffers[name] = value
            owner._is_stateful = batch_size is not None

    def enable_statefulness(self, batch_size: int) -> None:
        self._load_defaults(batch_size)

    def disable_statefulness(self) -> None:
        self._load_defaults(None)

    @contextmanager
    def statefulness(self, batch_size: int):
        self.enable_statefulness(batch_size)
        try:
            yield
        finally:
            self.disable_statefulness()


class ModuleList(nn.ModuleList, Module):
    pass


class ModuleDict(nn.ModuleDict, Module):
    pass
