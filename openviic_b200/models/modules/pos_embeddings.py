"""1-D sinusoid positional embedding over visual tokens (reference: pos_embeddings.py:39-72)."""

from __future__ import annotations

import math

import torch
from torch import nn

from ..utils import visual_position_table


class SinusoidPositionalEmbedding(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats, self.temperature, self.normalize = num_pos_feats, temperature, normalize
        self.scale = 2 * math.pi if scale is None else scale
        self._tables = {}

    def table(self, n: int, device) -> torch.Tensor:
        """(n, num_pos_feats) fp32 table for un-masked inputs: the embedding is input-independent."""
        key = (n, str(device))
        if key not in self._tables:
            self._tables[key] = visual_position_table(n, self.num_pos_feats, self.normalize, self.temperature).to(device)
        return self._tables[key]

    def forward(self, x, mask=None):
        if mask is None:
            return self.table(x.shape[1], x.device).unsqueeze(0).expand(x.shape[0], -1, -1)
        embed = (mask == False).cumsum(1, dtype=torch.float32)  # noqa: E712
        if self.normalize:
            embed = embed / (embed[:, -1:] + 1e-6) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=x.device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        pos = embed[:, :, None] / dim_t
        return torch.stack((pos[:, :, 0::2].sin(), pos[:, :, 1::2].cos()), dim=-1).flatten(-2)
