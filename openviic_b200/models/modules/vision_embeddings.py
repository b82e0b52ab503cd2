"""Visual feature embedding (reference: models/modules/vision_embeddings.py:8-20)."""

from __future__ import annotations

import torch
from torch import nn

from ... import ops
from ...builders.vision_embedding_builder import META_VISION_EMBEDDING


@META_VISION_EMBEDDING.register()
class FeatureEmbedding(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.proj = nn.Linear(config.D_FEATURE, config.D_MODEL)
        self.dropout = nn.Dropout(config.DROPOUT)

    def forward(self, features):
        """(B,n,D_FEATURE) -> (projected features fp32 (B,n,d), bool padding mask (B,1,1,n)); the mask
        comes from the RAW features (sum over the feature vector == 0)."""
        with torch.no_grad():
            feats16, mask = ops.feature_mask_cast(features)
            out = ops.linear(feats16, ops.cached_bf16(self.proj.weight), self.proj.bias, out_dtype=torch.float32)
            return out, mask.bool().unsqueeze(1).unsqueeze(1)
