"""Scripted-model known-answer test for the beam logic: V=6, eos=2, one image,
beam 3, max_len 4; step 0 prefers tokens 3,4,5 and every later step prefers <eos>."""

import torch


def scripted_logprobs(t: int, rows: int) -> torch.Tensor:
    lp = torch.full((rows, 1, 6), -5.0)
    if t == 0:
        lp[:, 0, 3], lp[:, 0, 4], lp[:, 0, 5] = -0.1, -0.2, -0.3
    else:
        lp[:, 0, 2] = -0.05
    return lp


def scripted_kat():
    ids = [[[3, 2, 0, 0], [4, 2, 0, 0], [5, 2, 0, 0]]]
    lp = [[[-0.1, -0.05, 0.0, 0.0], [-0.2, -0.05, 0.0, 0.0], [-0.3, -0.05, 0.0, 0.0]]]
    fed = [None, [3, 4, 5], [2, 2, 2], [0, 0, 0]]
    return ids, lp, fed


class ScriptedModel:
    """Anything with ``step(t, prev) -> (rows,1,V)`` and ``apply_to_states(fn)`` can be beam-searched."""

    def __init__(self, device="cpu"):
        self.device = torch.device(device)
        self.fed = []

    def step(self, t, prev):
        self.fed.append(None if prev is None else prev.reshape(-1).tolist())
        rows = 1 if t == 0 else prev.shape[0]
        return scripted_logprobs(t, rows).to(self.device)

    def apply_to_states(self, fn):
        pass
