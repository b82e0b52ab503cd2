"""Visual feature embeddings (reference: models/modules/vision_embeddings.py:8-71)."""

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


class _DualProjection(nn.Module):
    """region_proj / grid_proj shared by the two dual-path embeddings (vision_embeddings.py:22-31, 45-54)."""

    def __init__(self, config):
        super().__init__()
        self.region_proj = nn.Linear(config.D_REGION_FEATURE, config.D_MODEL)
        self.region_dropout = nn.Dropout(config.DROPOUT)
        self.grid_proj = nn.Linear(config.D_GRID_FEATURE, config.D_MODEL)
        self.grid_dropout = nn.Dropout(config.DROPOUT)

    def _project(self, region_features, grid_features):
        r16, r_mask = ops.feature_mask_cast(region_features)      # masks come from the RAW features
        g16, g_mask = ops.feature_mask_cast(grid_features)
        region = ops.linear(r16, ops.cached_bf16(self.region_proj.weight), self.region_proj.bias, out_dtype=torch.float32)
        grid = ops.linear(g16, ops.cached_bf16(self.grid_proj.weight), self.grid_proj.bias, out_dtype=torch.float32)
        return (region, r_mask.bool().unsqueeze(1).unsqueeze(1)), (grid, g_mask.bool().unsqueeze(1).unsqueeze(1))


@META_VISION_EMBEDDING.register()
class DualFeatureEmbedding(_DualProjection):
    """vision_embeddings.py:21-43."""

    def forward(self, region_features, grid_features):
        with torch.no_grad():
            return self._project(region_features, grid_features)


@META_VISION_EMBEDDING.register()
class GeometricDualFeatureEmbedding(_DualProjection):
    """vision_embeddings.py:45-71, with the two repairs without which it cannot run (documented in
    oracle/caption_oracle.py, P1 / P2): get_combine_masks' extra singleton dim is dropped, and the key-padding masks
    are expanded over the query dim before they are concatenated with the local region-to-grid masks."""

    def forward(self, region_features, region_boxes, grid_features, grid_boxes):
        with torch.no_grad():
            (region, r_mask), (grid, g_mask) = self._project(region_features, grid_features)
            n, g2 = region_features.shape[1], grid_features.shape[1]
            region2grid = ops.region_grid_mask(region_boxes, int(g2 ** 0.5))             # (B,1,n,g2)
            grid2region = region2grid.permute(0, 1, 3, 2)
            region2all = torch.cat([r_mask.expand(-1, -1, n, -1), region2grid], dim=-1)
            grid2all = torch.cat([grid2region, g_mask.expand(-1, -1, g2, -1)], dim=-1)
            return (region, r_mask), (grid, g_mask), (region2all, grid2all)
