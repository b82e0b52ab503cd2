"""The decode-state container (reference models/modules/containers.py:5-78): traversal order, batch expansion,
reset, and that states are ordinary buffers."""

import torch
from torch import nn

from openviic_b200.models.modules.containers import Module, ModuleList


class Leaf(Module):
    def __init__(self, tag):
        super().__init__()
        self.register_state("running", torch.zeros((0, 4)))
        self.register_state("count", torch.full((1,), tag, dtype=torch.long))
        self.register_state("maybe", None)


class Plain(nn.Module):          # not a container: neither it nor what it holds takes part
    def __init__(self):
        super().__init__()
        self.hidden = Leaf(99)


class Root(Module):
    def __init__(self):
        super().__init__()
        self.register_state("mask", torch.zeros((1, 1, 0), dtype=torch.bool))
        self.layers = ModuleList([Leaf(1), Leaf(2)])
        self.plain = Plain()
        self.last = Leaf(3)


def test_state_walk_expansion_and_reset():
    root = Root()
    assert [tuple(s.shape) if s is not None else None for s in root.states()] == \
        [(1, 1, 0)] + [(0, 4), (1,), None] * 3                 # own states, then container children depth-first
    assert "layers.1.count" in root.state_dict() and "plain.hidden.count" in root.state_dict()
    assert not root._is_stateful and not root.layers[0]._is_stateful

    with root.statefulness(5):
        assert root._is_stateful and root.layers[1]._is_stateful and root.last._is_stateful
        assert not root.plain.hidden._is_stateful              # below a plain nn.Module: untouched
        shapes = [tuple(s.shape) if s is not None else None for s in root.states()]
        assert shapes == [(5, 1, 1, 0)] + [(5, 0, 4), (5, 1), None] * 3
        assert [int(s[0, 0]) for s in root.states() if s is not None and s.dtype == torch.long] == [1, 2, 3]
        seen = []

        def grow(state):                                        # what the decoder and beam search do to states
            seen.append(None if state is None else tuple(state.shape))
            if state is None or state.dtype != torch.float32:
                return state
            return torch.cat([state, torch.ones(state.shape[0], 1, 4)], dim=1).repeat_interleave(2, dim=0)

        root.apply_to_states(grow)
        assert seen == shapes                                   # visited once each, in states() order
        assert root.layers[0].running.shape == (10, 1, 4) and root.last.running.shape == (10, 1, 4)
        assert root.plain.hidden.running.shape == (0, 4)

    assert not root._is_stateful and not root.last._is_stateful
    assert [tuple(s.shape) if s is not None else None for s in root.states()] == [(1, 1, 0)] + [(0, 4), (1,), None] * 3
    assert int(root.layers[1].count) == 2 and root.layers[0].maybe is None


def test_states_follow_the_module_device_and_dtype():
    root = Root().to(torch.float64)                             # buffers move with the module; defaults are re-cast on use
    root.enable_statefulness(2)
    assert root.last.running.shape == (2, 0, 4)
    root.disable_statefulness()
    assert root.last.running.shape == (0, 4)
