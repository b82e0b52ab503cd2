"""Stateful module container: the decode-state registry beam search reorders.

Same API as the reference's models/modules/containers.py:5-78 (``register_state``, ``states``,
``apply_to_states``, ``enable/disable_statefulness``, ``statefulness``).  States are registered
buffers, so they also appear in ``state_dict()`` exactly like the reference's.
"""

from __future__ import annotations

from contextlib import contextmanager

from torch import nn


class Module(nn.Module):
    def __init__(self):
        super().__init__()
        self._is_stateful = False
        self._state_names = []
        self._state_defaults = dict()

    def register_state(self, name: str, default):
        self._state_names.append(name)
        self._state_defaults[name] = None if default is None else default.clone().detach()
        self.register_buffer(name, default)

    def _stateful_children(self):
        return (m for m in self.children() if isinstance(m, Module))

    def states(self):
        for name in self._state_names:
            yield self._buffers[name]
        for child in self._stateful_children():
            yield from child.states()

    def apply_to_states(self, fn):
        for name in self._state_names:
            self._buffers[name] = fn(self._buffers[name])
        for child in self._stateful_children():
            child.apply_to_states(fn)

    def _fresh_state(self, name: str):
        default = self._state_defaults[name]
        if default is None:
            return None
        return default.clone().detach().to(self._buffers[name].device)

    def enable_statefulness(self, batch_size: int):
        for child in self._stateful_children():
            child.enable_statefulness(batch_size)
        for name in self._state_names:
            state = self._fresh_state(name)
            if state is not None:
                state = state.unsqueeze(0).expand([batch_size] + list(state.shape)).contiguous()
            self._buffers[name] = state
        self._is_stateful = True

    def disable_statefulness(self):
        for child in self._stateful_children():
            child.disable_statefulness()
        for name in self._state_names:
            self._buffers[name] = self._fresh_state(name)
        self._is_stateful = False

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
