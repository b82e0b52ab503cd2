"""Encoders (reference: models/modules/encoders.py:11-112), running on the sm_100a kernels."""

from __future__ import annotations

import torch
from torch import nn

from ... import ops
from ...builders.encoder_builder import META_ENCODER
from ..utils import clones
from .attentions import MultiHeadAttention
from .pos_embeddings import SinusoidPositionalEmbedding
from .positionwise_feed_forward import PositionWiseFeedForward


class EncoderLayer(nn.Module):
    """Attention -> feed-forward -> zero the padded query rows (encoders.py:11-22)."""

    def __init__(self, config):
        super().__init__()
        self.mhatt = MultiHeadAttention(config)
        self.pwff = PositionWiseFeedForward(config)

    def forward(self, queries, keys, values, padding_mask, attention_mask, **kwargs):
        """``padding_mask`` (B,1,1,nq) marks the padded QUERY rows that are zeroed after the feed-forward block."""
        att = self.mhatt(queries=queries, keys=keys, values=values, padding_mask=padding_mask,
                         attention_mask=attention_mask, **kwargs)
        return self.pwff(att, zero_rows=padding_mask.squeeze(1).squeeze(1))


class _LayerStack(nn.Module):
    """LN(x) + pos(x), then the layer stack; shared by the three single-stream encoders."""

    def __init__(self, config):
        super().__init__()
        self.pos_embedding = SinusoidPositionalEmbedding(config.D_MODEL)
        self.layer_norm = nn.LayerNorm(config.D_MODEL)
        self.d_model = config.D_MODEL
        self.layers = nn.ModuleList([EncoderLayer(config.SELF_ATTENTION) for _ in range(config.LAYERS)])

    def _embed(self, features):
        n = features.shape[1]
        return ops.add_layernorm(features, None, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps,
                                 pos=self.pos_embedding.table(n, features.device))

    def _run(self, features, padding_mask, **kwargs):
        with torch.no_grad():
            out = self._embed(features)
            outs = []
            for layer in self.layers:
                out = layer(queries=out, keys=out, values=out, padding_mask=padding_mask,
                            attention_mask=padding_mask, **kwargs)
                outs.append(out)
            return outs


@META_ENCODER.register()
class Encoder(_LayerStack):
    """encoders.py:24-40."""

    def forward(self, features: torch.Tensor, padding_mask: torch.Tensor):
        return self._run(features, padding_mask)[-1]


@META_ENCODER.register()
class MultilevelEncoder(_LayerStack):
    """Returns every layer's output stacked as (B, L, n, d) (encoders.py:42-63)."""

    def forward(self, features: torch.Tensor, padding_mask: torch.Tensor):
        return torch.stack(self._run(features, padding_mask), dim=1)


@META_ENCODER.register()
class GeometricEncoder(_LayerStack):
    """Box-relation biased encoder (encoders.py:65-112); takes the boxes next to the features."""

    def __init__(self, config):
        super().__init__(config)
        self.trignometric_embedding = config.TRIGNOMETRIC_EMBEDDING
        self.d_g = config.D_MODEL // config.SELF_ATTENTION.HEAD if self.trignometric_embedding else 4
        self.fc_gs = clones(nn.Linear(self.d_g, 1), config.SELF_ATTENTION.HEAD)
        for fc_g in self.fc_gs:
            nn.init.xavier_uniform_(fc_g.weight)
            nn.init.constant_(fc_g.bias, 0)

    def forward(self, features: torch.Tensor, boxes: torch.Tensor, padding_mask: torch.Tensor):
        with torch.no_grad():
            w_g = torch.cat([fc.weight for fc in self.fc_gs], dim=0)
            b_g = torch.cat([fc.bias for fc in self.fc_gs], dim=0)
            geometry = ops.geometry_bias(boxes, w_g, b_g, bool(self.trignometric_embedding))
        return self._run(features, padding_mask, relative_geometry_weights=geometry)[-1]


@META_ENCODER.register()
class CrossAttentionMultiLevelEncoder(_LayerStack):
    """CamoTransformer's encoder (encoders.py:213-249): the three layer outputs attend to each other through one extra
    attention block, and an MLP over the concatenation of the ORIGINAL three outputs is mixed in.  Kept as written:
    three layers are assumed (:235), and mlp1 reads the un-updated outputs (:242)."""

    def __init__(self, config):
        super().__init__(config)
        self.self_attn = MultiHeadAttention(config.SELF_ATTENTION)
        self.mlp1 = nn.Linear(3 * config.D_MODEL, config.D_MODEL)
        self.mlp2 = nn.Linear(config.D_MODEL, config.D_MODEL)

    def forward(self, features: torch.Tensor, padding_mask: torch.Tensor):
        with torch.no_grad():
            outs = self._run(features, padding_mask)
            out1, out2, out3 = outs
            out2 = 0.1 * self.self_attn(queries=out2, keys=out1, values=out1, padding_mask=padding_mask,
                                        attention_mask=padding_mask) + out2
            out3 = 0.1 * self.self_attn(queries=out3, keys=out2, values=out2, padding_mask=padding_mask,
                                        attention_mask=padding_mask) + out3
            mixed = ops.linear(torch.cat(outs, dim=-1), ops.cached_bf16(self.mlp1.weight), self.mlp1.bias,
                               act=ops.ACT_LEAKY_RELU, out_dtype=torch.float32)
            mixed = ops.linear(mixed, ops.cached_bf16(self.mlp2.weight), self.mlp2.bias, act=ops.ACT_LEAKY_RELU,
                               out_dtype=torch.float32)
            return out3 + 0.2 * mixed


@META_ENCODER.register()
class DualCollaborativeLevelEncoder(nn.Module):
    """Region and grid streams with geometry-biased self-attention, then locally-constrained cross-attention of each
    stream over the concatenation of both (encoders.py:114-211).  One repair (P3 in oracle/caption_oracle.py): the
    cross blocks zero the padded rows of their QUERY stream -- the reference hands them the (B,1,nq,nk) attention mask
    for that, which cannot work (encoders.py:197,205 vs :20)."""

    def __init__(self, config):
        super().__init__()
        self.d_model = config.D_MODEL
        self.trignometric_embedding = config.TRIGNOMETRIC_EMBEDDING
        self.d_g = config.D_MODEL // config.HEAD if self.trignometric_embedding else 4
        self.layer_norm_region = nn.LayerNorm(self.d_model)
        self.layer_norm_grid = nn.LayerNorm(self.d_model)
        self.fc_gs = clones(nn.Linear(self.d_g, 1), config.HEAD)
        self.pos_embedding = SinusoidPositionalEmbedding(config.D_MODEL, normalize=True)
        self.layers_region = nn.ModuleList([EncoderLayer(config.SELF_ATTENTION) for _ in range(config.LAYERS)])
        self.layers_grid = nn.ModuleList([EncoderLayer(config.SELF_ATTENTION) for _ in range(config.LAYERS)])
        self.region2grid = nn.ModuleList([EncoderLayer(config.CROSS_ATTENTION) for _ in range(config.LAYERS)])
        self.grid2region = nn.ModuleList([EncoderLayer(config.CROSS_ATTENTION) for _ in range(config.LAYERS)])
        for fc_g in self.fc_gs:
            nn.init.xavier_uniform_(fc_g.weight)
            nn.init.constant_(fc_g.bias, 0)

    def forward(self, region_features, region_boxes, region_padding_mask, region2all_mask,
                grid_features, grid_boxes, grid_padding_mask, grid2all_mask):
        with torch.no_grad():
            n = region_features.shape[1]
            dev = region_features.device
            boxes = torch.cat([region_boxes, grid_boxes], dim=1)
            w_g = torch.cat([fc.weight for fc in self.fc_gs], dim=0)
            b_g = torch.cat([fc.bias for fc in self.fc_gs], dim=0)
            geo = ops.geometry_bias(boxes, w_g, b_g, bool(self.trignometric_embedding))   # (B,h,n+g,n+g)
            ln_r, ln_g = self.layer_norm_region, self.layer_norm_grid
            region = ops.add_layernorm(region_features, None, ln_r.weight, ln_r.bias, ln_r.eps,
                                       pos=self.pos_embedding.table(n, dev))
            grid = ops.add_layernorm(grid_features, None, ln_g.weight, ln_g.bias, ln_g.eps,
                                     pos=self.pos_embedding.table(grid_features.shape[1], dev))
            for l_region, l_grid, l_r2g, l_g2r in zip(self.layers_region, self.layers_grid, self.region2grid, self.grid2region):
                region = l_region(queries=region, keys=region, values=region, padding_mask=region_padding_mask,
                                  attention_mask=region_padding_mask, relative_geometry_weights=geo[:, :, :n, :n])
                grid = l_grid(queries=grid, keys=grid, values=grid, padding_mask=grid_padding_mask,
                              attention_mask=grid_padding_mask, relative_geometry_weights=geo[:, :, n:, n:])
                combined = torch.cat([region, grid], dim=1)
                combined = combined + self.pos_embedding.table(combined.shape[1], dev)
                region = l_r2g(queries=region, keys=combined, values=combined, padding_mask=region_padding_mask,
                               attention_mask=region2all_mask, relative_geometry_weights=geo[:, :, :n, :])
                grid = l_g2r(queries=grid, keys=combined, values=combined, padding_mask=grid_padding_mask,
                             attention_mask=grid2all_mask, relative_geometry_weights=geo[:, :, n:, :])
            return torch.cat([region, grid], dim=1), torch.cat([region_padding_mask, grid_padding_mask], dim=-1)
