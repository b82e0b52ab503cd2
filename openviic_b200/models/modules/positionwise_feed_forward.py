"""Position-wise feed-forward block (reference: models/modules/positionwise_feed_forward.py:6-28)."""

from __future__ import annotations

import torch
from torch import nn

from ... import ops


class PositionWiseFeedForward(nn.Module):
    def __init__(self, config) -> None:
        super().__init__()
        self.fc1 = nn.Linear(config.D_MODEL, config.D_FF)
        self.fc2 = nn.Linear(config.D_FF, config.D_MODEL)
        self.dropout = nn.Dropout(p=config.DROPOUT)
        self.dropout_2 = nn.Dropout(p=config.DROPOUT)
        self.layer_norm = nn.LayerNorm(config.D_MODEL)

    def forward(self, input, zero_rows=None) -> torch.Tensor:
        """LN(x + fc2(relu(fc1 x))); ``zero_rows`` (bool/uint8 per row) fuses the callers' padded-row
        ``masked_fill(…, 0)`` (encoders.py:20, decoders.py:26) into the LayerNorm kernel."""
        with torch.no_grad():
            x = input
            hidden = ops.linear(x, ops.cached_bf16(self.fc1.weight), self.fc1.bias, act=ops.ACT_RELU)
            y = ops.linear(hidden, ops.cached_bf16(self.fc2.weight), self.fc2.bias, out_dtype=torch.float32)
            return ops.add_layernorm(y, x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps,
                                     zero_rows=zero_rows)
