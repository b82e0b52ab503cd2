"""Tables and masks used by the registered modules (reference: models/utils.py)."""

from __future__ import annotations

import copy
import math

import torch
from torch import nn


def sinusoid_encoding_table(max_len: int, d_model: int, padding_idx=None) -> torch.Tensor:
    """Decoder position table (models/utils.py:21-40): sin on even columns, cos on odd ones."""
    pos = torch.arange(max_len, dtype=torch.float32).view(-1, 1)
    half = torch.arange(d_model // 2, dtype=torch.float32).view(1, -1)
    angle = pos / 10000 ** (2 * half / d_model)
    table = torch.zeros(max_len, d_model)
    table[:, 0::2] = torch.sin(angle)
    table[:, 1::2] = torch.cos(angle)
    if padding_idx is not None:
        table[padding_idx] = 0
    return table


def visual_position_table(n: int, d_model: int, normalize: bool = False, temperature: float = 10000.0) -> torch.Tensor:
    """SinusoidPositionalEmbedding without a mask (pos_embeddings.py:58-72): positions 1..n."""
    embed = torch.arange(1, n + 1, dtype=torch.float32)
    if normalize:
        embed = embed / (embed[-1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(d_model, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / d_model)
    pos = embed[:, None] / dim_t
    return torch.stack((pos[:, 0::2].sin(), pos[:, 1::2].cos()), dim=-1).flatten(-2)


def generate_padding_mask(sequences, padding_idx: int):
    """(bs, seq_len[, dim]) -> bool (bs,1,1,seq_len), True where sum over dim == padding_idx
    (models/utils.py:48-61)."""
    if sequences is None:
        return None
    seq = sequences.unsqueeze(-1) if sequences.dim() == 2 else sequences
    return (torch.sum(seq, dim=-1) == padding_idx).unsqueeze(1).unsqueeze(1)


def generate_sequential_mask(seq_len: int) -> torch.Tensor:
    """Causal mask (1,1,seq_len,seq_len), True above the diagonal (models/utils.py:63-70)."""
    return torch.triu(torch.ones((seq_len, seq_len)), diagonal=1).to(torch.bool).unsqueeze(0).unsqueeze(0)


def clones(module: nn.Module, n: int) -> nn.ModuleList:
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])
