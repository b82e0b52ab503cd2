"""CPU fp32 oracle for OpenViIC's caption-generation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``openviic_b200/`` may import this file.  It is the
checker used by ``tests/``, by ``__graft_entry__.smoke()`` and by the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- never the thing that is shipped or measured as
the product.

What it is: a functional (no ``nn.Module``) restatement, in plain PyTorch fp32 on the CPU, of
the algorithm the reference executes for ``model.beam_search`` / ``model.forward``.  It keeps
the reference's operation order *as written* -- including the redundant work (cross-attention
K/V re-projected for every beam row at every step, self-attention K/V re-projected for every
cached token, every state gathered every step) -- so that

  * its results equal the reference's to fp32 round-off (pinned by ``tests/golden/*.npz``,
    generated from the real reference by ``oracle/ref_harness/gen_golden.py``), and
  * timing it is a fair stand-in for "the reference's CPU path" (``cpu_baseline.kind =
    "port"``), because ``/root/reference`` itself does not exist on the GPU box.

Parity status: PINNED.  The reference ships no tests or golden vectors,
so the pin is the reference itself, imported and run in the build container with fixed seeds.

All citations are ``file:line`` relative to the reference repository root.
Weights are addressed by the reference's own ``state_dict`` names.
"""

from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]

NEG_SENTINEL = -999.0  # models/modules/beam_search.py:54


# ----------------------------------------------------------------------------------------------
# Tables and masks
# ----------------------------------------------------------------------------------------------

def word_position_table(rows: int, d_model: int) -> Tensor:
    """Decoder position table: sin on even columns, cos on odd, row 0 zeroed.

    models/utils.py:21-40 (``positional_embedding`` + ``sinusoid_encoding_table`` with
    ``padding_idx=0``), consumed at models/modules/decoders.py:87-88.
    """
    pos = torch.arange(rows, dtype=torch.float32).view(-1, 1)
    j = torch.arange(d_model // 2, dtype=torch.float32).view(1, -1)
    angle = pos / 10000 ** (2 * j / d_model)
    table = torch.zeros(rows, d_model)
    table[:, 0::2] = torch.sin(angle)
    table[:, 1::2] = torch.cos(angle)
    table[0] = 0
    return table


def visual_position_table(n: int, d_model: int, normalize: bool = False) -> Tensor:
    """DETR-style 1-D sinusoid over token index 1..n (no mask => cumsum of ones).

    models/modules/pos_embeddings.py:58-72.  Input-independent, so a constant (n, d) table.
    """
    embed = torch.arange(1, n + 1, dtype=torch.float32)
    if normalize:
        embed = embed / (embed[-1:] + 1e-6) * (2 * math.pi)
    dim_t = torch.arange(d_model, dtype=torch.float32)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / d_model)
    pos = embed[:, None] / dim_t
    return torch.stack((pos[:, 0::2].sin(), pos[:, 1::2].cos()), dim=-1).flatten(-2)


def feature_padding_mask(x: Tensor) -> Tensor:
    """True where a visual token is padding: the sum over its raw feature vector is 0.

    models/utils.py:48-61 called from models/modules/vision_embeddings.py:16.  -> (B,1,1,n) bool.
    """
    return (x.sum(dim=-1) == 0)[:, None, None, :]


def box_relation_embedding(boxes: Tensor, dim_g: int, trig: bool, wave_len: float = 1000.0) -> Tensor:
    """Pairwise box-geometry features, models/utils.py:156-215.  boxes (B,n,4) -> (B,n,n,dim_g)."""
    b = boxes.size(0)
    x0, y0, x1, y1 = torch.chunk(boxes, 4, dim=-1)
    cx, cy = (x0 + x1) * 0.5, (y0 + y1) * 0.5
    w, h = (x1 - x0) + 1.0, (y1 - y0) + 1.0
    dx = torch.log(torch.clamp(torch.abs((cx - cx.view(b, 1, -1)) / w), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy - cy.view(b, 1, -1)) / h), min=1e-3))
    dw = torch.log(w / w.view(b, 1, -1))
    dh = torch.log(h / h.view(b, 1, -1))
    mat = torch.stack((dx, dy, dw, dh), dim=-1)  # (B,n,n,4)
    if not trig:
        return mat
    rng = torch.arange(dim_g / 8)
    freq = 1.0 / torch.pow(wave_len, rng / (dim_g / 8))
    scaled = (100.0 * mat).unsqueeze(-1) * freq.view(1, 1, 1, 1, -1)
    scaled = scaled.flatten(-2)
    return torch.cat((torch.sin(scaled), torch.cos(scaled)), dim=-1)


# ----------------------------------------------------------------------------------------------
# Attention variants (A1-A3), the multi-head wrapper (A5) and the feed-forward block (F1)
# ----------------------------------------------------------------------------------------------

# Precision model of the CUDA path (second checker, OFF by default -- the default IS the reference's fp32 arithmetic).
# With ``operand_rounding("bf16")`` active, every GEMM input is rounded to bf16 and the tensors the CUDA path stores
# as bf16 (q / k / v projections, the FFN hidden layer, scaled memory slots) are rounded where it stores them; sums,
# softmax, LayerNorm, the residual stream and the logits stay fp32, as on the GPU.  The reference algorithm is
# unchanged: this answers "does the CUDA path compute the reference's algorithm, given bf16 operands?" to ~1e-3,
# separately from "how far do bf16 operands move the result?" (measured against the fp32 default).
_OPERANDS: Optional[str] = None
_BF16_STORED = ("fc_q", "fc_k", "fc_v", "fc_s", "fc1")


class operand_rounding:
    """Context manager: ``with operand_rounding("bf16"): ...`` -- see the note above."""

    def __init__(self, mode: Optional[str]):
        assert mode in (None, "bf16")
        self.mode = mode

    def __enter__(self):
        global _OPERANDS
        self.prev, _OPERANDS = _OPERANDS, self.mode
        return self

    def __exit__(self, *exc):
        global _OPERANDS
        _OPERANDS = self.prev


def _rnd(x: Tensor) -> Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if _OPERANDS == "bf16" else x


# Dropout of the training step (T1).  The reference's nn.Dropout draws from torch's generator, which nothing outside
# torch can reproduce; the training restatement (and the patched reference of oracle/ref_harness/gen_golden_train.py, and
# the CUDA kernel cap_train_dropout) use a counter-based mask instead: element i of the tensor at the nn.Dropout module
# named `site` is kept iff hash(i, seed, crc32(site)) >= floor(p * 2^32); kept elements are scaled by float32(1 / (1 - p)).
_DROPOUT_SEED: Optional[int] = None   # None: eval semantics (dropout = identity)


class train_dropout:
    """Context manager: dropout active with this seed (one seed per optimizer step)."""

    def __init__(self, seed: Optional[int]):
        self.seed = seed

    def __enter__(self):
        global _DROPOUT_SEED
        self.prev, _DROPOUT_SEED = _DROPOUT_SEED, self.seed
        return self

    def __exit__(self, *exc):
        global _DROPOUT_SEED
        _DROPOUT_SEED = self.prev


def dropout_site(name: str) -> int:
    import zlib
    return zlib.crc32(name.encode())


def dropout_threshold(p: float) -> int:
    return int(float(p) * 4294967296.0)


def dropout_keep(numel: int, seed: int, site: str, p: float) -> Tensor:
    """bool (numel,): the mask of cap_train_dropout (csrc/train.cu dropout_hash), in 32-bit wrap-around arithmetic."""
    import numpy as np
    m = np.uint64(0xFFFFFFFF)
    x = (np.arange(numel, dtype=np.uint64) * np.uint64(0x9E3779B1) + np.uint64(seed & 0xFFFFFFFF) * np.uint64(0x85EBCA77)
         + np.uint64(dropout_site(site)) * np.uint64(0xC2B2AE3D)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x2C1B3C6D)) & m
    x ^= x >> np.uint64(12)
    x = (x * np.uint64(0x297A2D39)) & m
    x ^= x >> np.uint64(15)
    return torch.from_numpy(x >= np.uint64(dropout_threshold(p)))


def _drop(x: Tensor, site: str, p: float) -> Tensor:
    if _DROPOUT_SEED is None or not p:
        return x
    keep = dropout_keep(x.numel(), _DROPOUT_SEED, site, p).view(x.shape)
    return x * keep * torch.tensor(1.0 / (1.0 - float(p)), dtype=torch.float32)


def _lin(w: Weights, name: str, x: Tensor) -> Tensor:
    y = F.linear(_rnd(x), w[name + ".weight"], w.get(name + ".bias"))
    return _rnd(y) if name.rsplit(".", 1)[-1] in _BF16_STORED else y


def dot_product_attention(w: Weights, p: str, att_cfg, queries: Tensor, keys: Tensor, values: Tensor,
                          mask: Optional[Tensor], geometry: Optional[Tensor] = None) -> Tensor:
    """The three live attention variants, selected by ``att_cfg.ARCHITECTURE``.

    ScaledDotProductAttention                    models/modules/attentions.py:44-58
    AugmentedGeometryScaledDotProductAttention   models/modules/attentions.py:97-114
    AugmentedMemoryScaledDotProductAttention     models/modules/attentions.py:158-185
    """
    kind = att_cfg.ARCHITECTURE
    h, d_k, d_v = att_cfg.HEAD, att_cfg.D_KEY, att_cfg.D_VALUE
    b, nq = queries.shape[:2]
    nk = keys.shape[1]
    q = _lin(w, p + "fc_q", queries).view(b, nq, h, d_k).permute(0, 2, 1, 3)
    k = _lin(w, p + "fc_k", keys)
    v = _lin(w, p + "fc_v", values)
    if kind == "AugmentedMemoryScaledDotProductAttention":
        m = att_cfg.MEMORY
        k = torch.cat([k, _rnd(math.sqrt(d_k) * w[p + "m_k"]).expand(b, m, h * d_k)], 1)
        v = torch.cat([v, _rnd(math.sqrt(m) * w[p + "m_v"]).expand(b, m, h * d_v)], 1)
    nk_all = k.shape[1]
    k = k.view(b, nk_all, h, d_k).permute(0, 2, 3, 1)
    v = v.view(b, nk_all, h, d_v).permute(0, 2, 1, 3)
    att = torch.matmul(q, k) / math.sqrt(d_k)
    if mask is not None:
        if kind == "AugmentedMemoryScaledDotProductAttention":
            att[:, :, :, :nk] = att[:, :, :, :nk].masked_fill(mask, -math.inf)  # memory slots never masked
        else:
            att = att.masked_fill(mask, -math.inf)
    if kind == "AugmentedGeometryScaledDotProductAttention":
        att = torch.log(torch.clamp(geometry, min=1e-6)) + att
    att = torch.softmax(att, dim=-1)
    out = torch.matmul(att, v).permute(0, 2, 1, 3).contiguous().view(b, nq, h * d_v)
    return _lin(w, p + "fc_o", out)


def adaptive_attention(w: Weights, p: str, att_cfg, queries: Tensor, keys: Tensor, values: Tensor, signals: Tensor,
                       mask: Optional[Tensor]) -> Tensor:
    """AdaptiveScaledDotProductAttention.forward, models/modules/attentions.py:229-268: query i gets one extra softmax
    column q_i . s_i / sqrt(d_k) (the DIAGONAL of q s^T, :251-252) whose value is s_i (:257).  Vectorised, same
    arithmetic.  Pinned at the operator level (oracle/ref_harness/gen_golden_ops.py): the only caller in the
    reference, AdaptiveDecoder, cannot be constructed."""
    h, d_k, d_v = att_cfg.HEAD, att_cfg.D_KEY, att_cfg.D_VALUE
    b, nq = queries.shape[:2]
    nk = keys.shape[1]
    q = _lin(w, p + "fc_q", queries).view(b, nq, h, d_k).permute(0, 2, 1, 3)
    s = _lin(w, p + "fc_s", signals).view(b, nq, h, d_k).permute(0, 2, 1, 3)          # fc_s is bf16-stored like fc_q
    k = _lin(w, p + "fc_k", keys).view(b, nk, h, d_k).permute(0, 2, 3, 1)
    v = _lin(w, p + "fc_v", values).view(b, nk, h, d_v).permute(0, 2, 1, 3)
    att = torch.matmul(q, k) / math.sqrt(d_k)
    if mask is not None:
        att = att.masked_fill(mask, -math.inf)
    lang = (q * s).sum(-1, keepdim=True) / math.sqrt(d_k)                              # (b,h,nq,1)
    prob = torch.softmax(torch.cat([att, lang], dim=-1), dim=-1)
    out = torch.matmul(prob[..., :nk], v) + prob[..., nk:] * s
    return _lin(w, p + "fc_o", out.permute(0, 2, 1, 3).contiguous().view(b, nq, h * d_v))


def multi_head_attention(w: Weights, p: str, att_cfg, queries: Tensor, keys: Tensor, values: Tensor,
                         mask: Optional[Tensor], cache: Optional[dict] = None,
                         geometry: Optional[Tensor] = None) -> Tensor:
    """MultiHeadAttention.forward, models/modules/attentions.py:296-317 (eval: dropout = identity).

    ``cache`` (stateful decode) holds the raw, pre-projection inputs, exactly like the reference.
    """
    if cache is not None:
        cache["keys"] = torch.cat([cache["keys"], keys], 1)
        cache["values"] = torch.cat([cache["values"], values], 1)
        keys, values = cache["keys"], cache["values"]
    out = dot_product_attention(w, p + "attention.", att_cfg, queries, keys, values, mask, geometry)
    if _DROPOUT_SEED is not None:
        out = _drop(out, p + "dropout", att_cfg.DROPOUT)        # attentions.py:308
    d = queries.shape[-1]
    out = F.layer_norm(queries + out, (d,), w[p + "layer_norm.weight"], w[p + "layer_norm.bias"])
    if att_cfg.USE_AOA:
        x = torch.cat([queries, out], dim=-1)
        out = _lin(w, p + "informative_attention", x) * torch.sigmoid(_lin(w, p + "gated_attention", x))
    return out


def feed_forward(w: Weights, p: str, x: Tensor, p_drop: float = 0.0) -> Tensor:
    """PositionWiseFeedForward.forward, models/modules/positionwise_feed_forward.py:23-28."""
    y = _lin(w, p + "fc2", _drop(F.relu(_lin(w, p + "fc1", x)), p + "dropout_2", p_drop))
    y = _drop(y, p + "dropout", p_drop)
    return F.layer_norm(x + y, (x.shape[-1],), w[p + "layer_norm.weight"], w[p + "layer_norm.bias"])


# ----------------------------------------------------------------------------------------------
# Encoders (E0-E3)
# ----------------------------------------------------------------------------------------------

def _encoder_layer(w: Weights, p: str, att_cfg, x: Tensor, pad_mask: Tensor, geometry=None) -> Tensor:
    """EncoderLayer.forward, models/modules/encoders.py:17-22."""
    att = multi_head_attention(w, p + "mhatt.", att_cfg, x, x, x, pad_mask, geometry=geometry)
    ff = feed_forward(w, p + "pwff.", att, att_cfg.DROPOUT if _DROPOUT_SEED is not None else 0.0)
    return ff.masked_fill(pad_mask.squeeze(1).squeeze(1).unsqueeze(-1), 0)


def geometry_weights(w: Weights, p: str, enc_cfg, boxes: Tensor) -> Tensor:
    """Per-head ReLU(Linear(box embedding)), models/modules/encoders.py:94-101.  -> (B,h,n,n)."""
    h = enc_cfg.SELF_ATTENTION.HEAD
    trig = bool(enc_cfg.TRIGNOMETRIC_EMBEDDING)
    d_g = enc_cfg.D_MODEL // h if trig else 4
    emb = box_relation_embedding(boxes, d_g, trig)
    b, n = emb.shape[:2]
    flat = emb.view(-1, d_g)
    per_head = [_lin(w, f"{p}fc_gs.{i}", flat).view(b, 1, n, n) for i in range(h)]
    return F.relu(torch.cat(per_head, dim=1))


def encode(w: Weights, model_cfg, feats: Tensor, boxes: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """``encoder_forward``: vision embedding + encoder stack.

    FeatureEmbedding      models/modules/vision_embeddings.py:15-20
    Encoder               models/modules/encoders.py:35-40
    MultilevelEncoder     models/modules/encoders.py:53-63   (returns all levels, (B,L,n,d))
    CrossAttentionMultiLevelEncoder  models/modules/encoders.py:213-249 (CamoTransformer)
    GeometricEncoder      models/modules/encoders.py:93-112  (called with boxes -- the documented
                          ORT call-site patch)
    """
    enc_cfg = model_cfg.ENCODER
    if enc_cfg.ARCHITECTURE == "DualCollaborativeLevelEncoder":   # feats = (region, grid), boxes = (region boxes, grid boxes)
        (region, region_mask), (grid, grid_mask), (region2all, grid2all) = dual_feature_embedding(w, feats[0], boxes[0], feats[1], boxes[1])
        return dual_collaborative_encode(w, enc_cfg, region, boxes[0], region_mask, region2all, grid, boxes[1], grid_mask, grid2all)
    pad_mask = feature_padding_mask(feats)
    x = _lin(w, "vision_embedding.proj", feats)
    if _DROPOUT_SEED is not None:
        x = _drop(x, "vision_embedding.dropout", model_cfg.VISION_EMBEDDING.DROPOUT)   # vision_embeddings.py:18
    d = enc_cfg.D_MODEL
    kind = enc_cfg.ARCHITECTURE
    geometry = geometry_weights(w, "encoder.", enc_cfg, boxes) if kind == "GeometricEncoder" else None
    x = F.layer_norm(x, (d,), w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"])
    x = x + visual_position_table(x.shape[1], d)
    levels = []
    for i in range(enc_cfg.LAYERS):
        x = _encoder_layer(w, f"encoder.layers.{i}.", enc_cfg.SELF_ATTENTION, x, pad_mask, geometry)
        levels.append(x)
    if kind == "MultilevelEncoder":
        return torch.stack(levels, dim=1), pad_mask
    if kind in ("Encoder", "GeometricEncoder"):
        return x, pad_mask
    if kind == "CrossAttentionMultiLevelEncoder":
        # models/modules/encoders.py:226-249: the layer outputs attend to each other through ONE extra attention block,
        # an MLP over the concatenation of the ORIGINAL three outputs is mixed in (hard-coded three layers, :235; the
        # un-updated `outs` feed mlp1, :242 -- both kept as written).
        att_cfg = enc_cfg.SELF_ATTENTION
        out1, out2, out3 = levels
        out2 = 0.1 * multi_head_attention(w, "encoder.self_attn.", att_cfg, out2, out1, out1, pad_mask) + out2
        out3 = 0.1 * multi_head_attention(w, "encoder.self_attn.", att_cfg, out3, out2, out2, pad_mask) + out3
        mixed = F.leaky_relu(_lin(w, "encoder.mlp1", torch.cat(levels, dim=-1)))
        mixed = F.leaky_relu(_lin(w, "encoder.mlp2", mixed))
        return out3 + 0.2 * mixed, pad_mask
    raise KeyError(f"oracle has no encoder {kind!r}")


# ----------------------------------------------------------------------------------------------
# Dual-path (region + grid) encoder, E4: GeometricDualFeatureEmbedding + DualCollaborativeLevelEncoder
# ----------------------------------------------------------------------------------------------
# The reference's dual-path pieces do not run as shipped.  This restatement -- and the patched
# reference composition in oracle/ref_harness/gen_golden_dlct.py that pins it -- apply exactly three repairs, each the
# only shape-consistent reading of the code:
#   P1  get_combine_masks returns (B,1,1,n,g*g) (models/utils.py:154: two unsqueeze(1)); one of the two singleton
#       dims is dropped so that .permute(0,1,3,2) (vision_embeddings.py:61) applies;
#   P2  the key-padding masks (B,1,1,n) are expanded over the query dim before torch.cat with the (B,1,n,g*g) local
#       masks (vision_embeddings.py:62-63; cat does not broadcast);
#   P3  the two cross blocks hand the (B,1,nq,nk) ATTENTION mask to EncoderLayer as `padding_mask`
#       (encoders.py:197,205), which uses it to zero padded QUERY rows (encoders.py:20): the query stream's own
#       padding mask is used for that instead.
# Nothing else departs from the reference.

def grid_cells_under_box(boxes: Tensor, grid_size: int) -> Tensor:
    """get_combine_masks / get_grids_by_corner / lower_bound, models/utils.py:100-154: True = the grid cell is NOT
    covered by the box's corner-to-corner cell rectangle.  boxes (B,n,4) in [0,1] -> (B,1,n,g*g) bool (P1 applied).

    lower_bound(nums, t) returns the last index with nums[i] <= t (0 when none is): cells are i/grid_size."""
    edges = torch.arange(grid_size, dtype=torch.float64) / grid_size

    def last_le(v):                      # (B,n) -> index of the last edge <= v, 0 if none
        le = edges.view(1, 1, -1) <= v.double().unsqueeze(-1)
        idx = le.long().cumsum(-1).argmax(-1)      # position of the last True when the Trues are a prefix
        return torch.where(le.any(-1), idx, torch.zeros_like(idx))

    x1, y1 = last_le(boxes[..., 0]), last_le(boxes[..., 1])
    x2, y3 = last_le(boxes[..., 2]), last_le(boxes[..., 3])
    top_left, top_right, bot_left = y1 * grid_size + x1, y1 * grid_size + x2, y3 * grid_size + x1
    width = top_right - top_left + 1                               # (B,n)
    cell = torch.arange(grid_size * grid_size).view(1, 1, -1)
    # rows start at top_left, top_left + g, ... <= bot_left; each covers [start, start + width)
    rel = cell - top_left.unsqueeze(-1)
    row, col = torch.div(rel, grid_size, rounding_mode="floor"), rel % grid_size
    covered = (rel >= 0) & (row * grid_size + top_left.unsqueeze(-1) <= bot_left.unsqueeze(-1)) & (col < width.unsqueeze(-1))
    # res[i:i+width] slices clip at the end of the array, and a slice may run past the row's end into the next row
    return (~covered).unsqueeze(1)


def dual_feature_embedding(w: Weights, region_feats: Tensor, region_boxes: Tensor, grid_feats: Tensor, grid_boxes: Tensor):
    """GeometricDualFeatureEmbedding.forward, models/modules/vision_embeddings.py:45-71 (eval: dropout = identity)."""
    region_mask, grid_mask = feature_padding_mask(region_feats), feature_padding_mask(grid_feats)
    n, g2 = region_feats.shape[1], grid_feats.shape[1]
    region2grid = grid_cells_under_box(region_boxes, int(g2 ** 0.5))                       # (B,1,n,g2)   P1
    grid2region = region2grid.permute(0, 1, 3, 2)                                          # (B,1,g2,n)
    region2all = torch.cat([region_mask.expand(-1, -1, n, -1), region2grid], dim=-1)        # (B,1,n,n+g2) P2
    grid2all = torch.cat([grid2region, grid_mask.expand(-1, -1, g2, -1)], dim=-1)           # (B,1,g2,n+g2)
    region = _lin(w, "vision_embedding.region_proj", region_feats)
    grid = _lin(w, "vision_embedding.grid_proj", grid_feats)
    return (region, region_mask), (grid, grid_mask), (region2all, grid2all)


def _dual_layer(w: Weights, p: str, att_cfg, q: Tensor, kv: Tensor, att_mask: Tensor, row_mask: Tensor, geometry: Tensor) -> Tensor:
    """EncoderLayer.forward (encoders.py:17-22) with separate attention and row masks (P3)."""
    att = multi_head_attention(w, p + "mhatt.", att_cfg, q, kv, kv, att_mask, geometry=geometry)
    ff = feed_forward(w, p + "pwff.", att)
    return ff.masked_fill(row_mask.squeeze(1).squeeze(1).unsqueeze(-1), 0)


def dual_collaborative_encode(w: Weights, enc_cfg, region, region_boxes, region_mask, region2all, grid, grid_boxes, grid_mask,
                              grid2all) -> Tuple[Tensor, Tensor]:
    """DualCollaborativeLevelEncoder.forward, models/modules/encoders.py:156-211."""
    h, d = enc_cfg.HEAD, enc_cfg.D_MODEL
    trig = bool(enc_cfg.TRIGNOMETRIC_EMBEDDING)
    d_g = d // h if trig else 4
    n = region.shape[1]
    boxes = torch.cat([region_boxes, grid_boxes], dim=1)
    emb = box_relation_embedding(boxes, d_g, trig)
    b, nk = emb.shape[:2]
    flat = emb.view(-1, d_g)
    geo = F.relu(torch.cat([_lin(w, f"encoder.fc_gs.{i}", flat).view(b, 1, nk, nk) for i in range(h)], dim=1))

    def pos(x):   # SinusoidPositionalEmbedding(d, normalize=True): the second assignment wins (encoders.py:121,135)
        return visual_position_table(x.shape[1], d, normalize=True)

    region = F.layer_norm(region, (d,), w["encoder.layer_norm_region.weight"], w["encoder.layer_norm_region.bias"]) + pos(region)
    grid = F.layer_norm(grid, (d,), w["encoder.layer_norm_grid.weight"], w["encoder.layer_norm_grid.bias"]) + pos(grid)
    for i in range(enc_cfg.LAYERS):
        region = _dual_layer(w, f"encoder.layers_region.{i}.", enc_cfg.SELF_ATTENTION, region, region, region_mask, region_mask,
                             geo[:, :, :n, :n])
        grid = _dual_layer(w, f"encoder.layers_grid.{i}.", enc_cfg.SELF_ATTENTION, grid, grid, grid_mask, grid_mask, geo[:, :, n:, n:])
        combined = torch.cat([region, grid], dim=1)
        combined = combined + pos(combined)
        region = _dual_layer(w, f"encoder.region2grid.{i}.", enc_cfg.CROSS_ATTENTION, region, combined, region2all, region_mask,
                             geo[:, :, :n, :])
        grid = _dual_layer(w, f"encoder.grid2region.{i}.", enc_cfg.CROSS_ATTENTION, grid, combined, grid2all, grid_mask,
                           geo[:, :, n:, :])
    return torch.cat([region, grid], dim=1), torch.cat([region_mask, grid_mask], dim=-1)


# ----------------------------------------------------------------------------------------------
# Decoders (D1-D3)
# ----------------------------------------------------------------------------------------------

def new_decode_state(model_cfg, rows: int) -> dict:
    """Registered states after ``enable_statefulness`` (models/modules/containers.py:34-56):
    empty running mask (R,1,1,0), running_seq (R,1)=0, empty per-layer raw K/V caches (R,0,d)."""
    d = model_cfg.DECODER.D_MODEL
    return {
        "mask": torch.zeros(rows, 1, 1, 0, dtype=torch.bool),
        "seq": torch.zeros(rows, 1, dtype=torch.long),
        "layers": [{"keys": torch.zeros(rows, 0, d), "values": torch.zeros(rows, 0, d)}
                   for _ in range(model_cfg.DECODER.LAYERS)],
    }


def _decoder_layer(w: Weights, p: str, dec_cfg, x: Tensor, enc: Tensor, pad_rows: Tensor,
                   self_mask: Tensor, enc_mask: Tensor, cache: Optional[dict]) -> Tensor:
    """DecoderLayer.forward models/modules/decoders.py:21-28; MeshedDecoderLayer.forward :51-73."""
    att_cfg = dec_cfg.ATTENTION
    s = multi_head_attention(w, p + "self_attn.", att_cfg.SELF_ATTENTION, x, x, x, self_mask, cache=cache)
    if dec_cfg.ARCHITECTURE == "MeshedDecoder":
        n_lv = att_cfg.N_ENCODER_LAYERS
        crosses = [multi_head_attention(w, p + "enc_attn.", att_cfg.ENC_ATTENTION, s, enc[:, i], enc[:, i], enc_mask)
                   for i in range(n_lv)]
        alphas = [torch.sigmoid(_lin(w, f"{p}fc_alphas.{i}", torch.cat([s, c], dim=-1)))
                  for i, c in enumerate(crosses)]
        mixed = 0
        for a, c in zip(alphas, crosses):
            mixed = mixed + a * c
        c = mixed / n_lv ** 0.5
    else:
        c = multi_head_attention(w, p + "enc_attn.", att_cfg.ENC_ATTENTION, s, enc, enc, enc_mask)
    # PositionWiseFeedForward(config.ENC_ATTENTION), decoders.py:19
    ff = feed_forward(w, p + "pwff.", c, att_cfg.ENC_ATTENTION.DROPOUT if _DROPOUT_SEED is not None else 0.0)
    return ff.masked_fill(pad_rows.unsqueeze(-1), 0)


def decode(w: Weights, model_cfg, tokens: Tensor, enc: Tensor, enc_mask: Tensor, pad_idx: int,
           state: Optional[dict] = None) -> Tensor:
    """Decoder.forward / MeshedDecoder.forward, models/modules/decoders.py:95-123 / :145-173.

    tokens (R,S) int64 -> log-probs (R,S,V).  With ``state`` it is the stateful single-step path.
    """
    dec_cfg = model_cfg.DECODER
    rows, s_len = tokens.shape
    pad = (tokens == pad_idx)                                   # (R,S)
    pad4 = pad[:, None, None, :]                                # generate_padding_mask on ids
    causal = torch.triu(torch.ones(s_len, s_len), diagonal=1).to(torch.bool)[None, None]
    self_mask = torch.logical_or(pad4, causal)
    if state is not None:
        state["mask"] = torch.cat([state["mask"], self_mask], -1)
        self_mask = state["mask"]
        state["seq"] = state["seq"] + 1                         # running_seq.add_(1); NOT zeroed for pad
        seq = state["seq"]
    else:
        seq = torch.arange(1, s_len + 1).view(1, -1).expand(rows, -1).masked_fill(pad, 0)
    x = F.embedding(tokens, w["decoder.word_emb.components.weight"]) + F.embedding(seq, w["decoder.pos_emb.weight"])
    for i in range(dec_cfg.LAYERS):
        cache = state["layers"][i] if state is not None else None
        x = _decoder_layer(w, f"decoder.layers.{i}.", dec_cfg, x, enc, pad, self_mask, enc_mask, cache)
    return F.log_softmax(F.linear(_rnd(x), w["decoder.fc.weight"]), dim=-1)


# ----------------------------------------------------------------------------------------------
# Beam search (B1-B4, S1)
# ----------------------------------------------------------------------------------------------

def _reorder(s: Tensor, b_s: int, cur: int, beam_idx: Tensor) -> Tensor:
    """BeamSearch._expand_state, models/modules/beam_search.py:19-34: view (B,cur,...) and pick
    ``beam_idx`` (B,beam) along dim 1, flatten back to (B*beam, ...)."""
    tail = list(s.shape[1:])
    s = s.view(b_s, cur, *tail)
    picked = s[torch.arange(b_s)[:, None], beam_idx]
    return picked.reshape(-1, *tail)


def beam_search(step: Callable[[int, Optional[Tensor]], Tensor], reorder_states: Callable[[Callable], None],
                b_s: int, beam: int, max_len: int, eos_idx: int, out_size: int = 1,
                trace: Optional[list] = None, stable_ties: bool = False, return_probs: bool = False):
    """BeamSearch.apply/iter/select, models/modules/beam_search.py:36-118.

    ``step(t, prev_tokens)`` returns (rows,1,V) log-probs; ``reorder_states(fn)`` maps ``fn`` over
    every decode state (models/modules/containers.py:27-32).  Always runs ``max_len`` steps.

    Tie order.  The reference calls ``torch.sort(descending=True)`` WITHOUT ``stable=True``.  Probed in
    the build container (torch 2.11 CPU): for rows of more than 16 elements that is an unstable
    introsort whose order among exactly equal candidates is implementation-defined (neither lowest-
    nor highest-index-first; the CUDA sort differs again), so the reference has no portable tie
    order to reproduce.  ``stable_ties=False`` restates the call as written (identical to the
    reference whenever the selected candidates are distinct); ``stable_ties=True`` pins the order
    this repo defines -- value descending, then lowest flat index -- which is what the CUDA
    kernels implement and what the tie-laden known-answer tests use.
    """
    seq_mask = torch.ones(b_s, beam, 1)
    seq_logprob = torch.zeros(b_s, 1, 1)
    outputs: List[Tensor] = []
    log_probs: List[Tensor] = []
    all_log_probs: List[Tensor] = []   # return_probs: beam_search.py:68-81 -- never reordered by later selections
    selected_words: Optional[Tensor] = None
    for t in range(max_len):
        cur = 1 if t == 0 else beam
        word_lp = step(t, selected_words).view(b_s, cur, -1)
        vocab = word_lp.shape[-1]
        cand = seq_logprob + word_lp
        if t > 0:
            alive = (selected_words.view(b_s, cur) != eos_idx).float().unsqueeze(-1)
            seq_mask = seq_mask * alive
            word_lp = word_lp * seq_mask.expand_as(word_lp)
            frozen = seq_logprob.expand_as(cand).contiguous()
            frozen[:, :, 1:] = NEG_SENTINEL
            cand = seq_mask * cand + frozen * (1 - seq_mask)
        # select(): full descending sort of the (cur*V) candidates, keep the first `beam`
        if stable_ties:
            sorted_lp, sorted_idx = torch.sort(cand.view(b_s, -1), stable=True, dim=-1, descending=True)
        else:
            sorted_lp, sorted_idx = torch.sort(cand.view(b_s, -1), -1, descending=True)
        sel_lp, sel_idx = sorted_lp[:, :beam], sorted_idx[:, :beam]
        sel_beam = torch.div(sel_idx, vocab, rounding_mode="trunc")
        sel_word = sel_idx - sel_beam * vocab
        reorder_states(lambda s, _b=sel_beam, _c=cur: _reorder(s, b_s, _c, _b))
        seq_logprob = sel_lp.unsqueeze(-1)
        seq_mask = torch.gather(seq_mask, 1, sel_beam.unsqueeze(-1))
        outputs = [torch.gather(o, 1, sel_beam.unsqueeze(-1)) for o in outputs]
        outputs.append(sel_word.unsqueeze(-1))
        if return_probs:
            all_log_probs.append((word_lp.expand(b_s, beam, -1) if t == 0 else word_lp).unsqueeze(2))
        this_lp = torch.gather(word_lp, 1, sel_beam.unsqueeze(-1).expand(b_s, beam, vocab))
        this_lp = torch.gather(this_lp, 2, sel_word.unsqueeze(-1))
        log_probs = [torch.gather(o, 1, sel_beam.unsqueeze(-1)) for o in log_probs]
        log_probs.append(this_lp)
        selected_words = sel_word.reshape(-1, 1)
        if trace is not None:
            trace.append({"beam": sel_beam.clone(), "word": sel_word.clone(), "seq_logprob": sel_lp.clone()})
    _, order = torch.sort(seq_logprob, 1, descending=True)
    ids = torch.gather(torch.cat(outputs, -1), 1, order.expand(b_s, beam, max_len))
    lps = torch.gather(torch.cat(log_probs, -1), 1, order.expand(b_s, beam, max_len))
    ids, lps = ids.contiguous()[:, :out_size], lps.contiguous()[:, :out_size]
    if out_size == 1:
        ids, lps = ids.squeeze(1), lps.squeeze(1)
    if return_probs:   # beam_search.py:103-118: gathered by the final order only, all `beam` rows kept
        probs = torch.cat(all_log_probs, 2)
        probs = torch.gather(probs, 1, order.unsqueeze(-1).expand(b_s, beam, max_len, probs.shape[-1]))
        return ids, lps, probs
    return ids, lps


# ----------------------------------------------------------------------------------------------
# Whole-path entry points (M1, M2)
# ----------------------------------------------------------------------------------------------

def caption_beam_search(w: Weights, model_cfg, vocab, feats: Tensor, boxes: Optional[Tensor] = None,
                        beam: int = 5, out_size: int = 1, trace: Optional[list] = None,
                        logits_trace: Optional[list] = None, return_probs: bool = False):
    """BaseTransformer.beam_search + .step, models/base_transformer.py:30-53.

    Returns (ids (B,T) int64, log_probs (B,T) fp32) for out_size == 1, else (B,out_size,T).
    """
    b_s = (feats[0] if isinstance(feats, (tuple, list)) else feats).shape[0]
    with torch.no_grad():
        enc, enc_mask = encode(w, model_cfg, feats, boxes)
        st = {"enc": enc, "enc_mask": enc_mask, "dec": new_decode_state(model_cfg, b_s)}

        def step(t, prev):
            rows = st["enc"].shape[0]
            tokens = torch.full((rows, 1), vocab.bos_idx, dtype=torch.long) if t == 0 else prev
            lp = decode(w, model_cfg, tokens, st["enc"], st["enc_mask"], vocab.padding_idx, st["dec"])
            if logits_trace is not None:
                logits_trace.append(lp.squeeze(1).clone())
            return lp

        def reorder_states(fn):
            # traversal order of containers.Module.apply_to_states
            st["enc"], st["enc_mask"] = fn(st["enc"]), fn(st["enc_mask"])
            d = st["dec"]
            d["mask"], d["seq"] = fn(d["mask"]), fn(d["seq"])
            for lay in d["layers"]:
                lay["keys"], lay["values"] = fn(lay["keys"]), fn(lay["values"])

        return beam_search(step, reorder_states, b_s, beam, vocab.max_caption_length, vocab.eos_idx,
                           out_size, trace, return_probs=return_probs)


def teacher_forced_log_probs(w: Weights, model_cfg, vocab, feats: Tensor, tokens: Tensor,
                             boxes: Optional[Tensor] = None) -> Tensor:
    """``model.forward(items)``: models/standard_stransformer.py:21-31 -> (B,T,V) log-probs."""
    with torch.no_grad():
        enc, enc_mask = encode(w, model_cfg, feats, boxes)
        return decode(w, model_cfg, tokens, enc, enc_mask, vocab.padding_idx, None)


# ----------------------------------------------------------------------------------------------
# T1: the XE training step (trainers/vi_trainer.py:100-119, trainers/base_trainer.py:87-91, 114-117)
# ----------------------------------------------------------------------------------------------
# The reference's step is: out = model(items) (teacher forcing, log-probs), NLLLoss(ignore_index=<pad>) over
# (B*T, V) against the shifted-right tokens, backward, Adam(lr, betas=(0.9, 0.98)) step, LambdaLR step with
# lambda(s) = d_model^-0.5 * min((s+1)^-0.5, (s+1) * warmup^-1.5).  Dropout: the reference's nn.Dropout modules draw from
# torch's generator, which has no portable restatement; `dropout_seeds` (one per step) switches the counter-based masks
# of _drop() on instead, None is the p = 0 computation.  oracle/ref_harness/gen_golden_train.py pins BOTH to the real
# reference's modules: with every nn.Dropout set to p = 0, and with every nn.Dropout's forward wrapped to apply the
# same counter-based mask at the same site name.
# nn.Embedding(padding_idx=<pad>) never updates the <pad> row: its gradient is zeroed here the same way; the
# sinusoid position table is frozen (decoders.py:87-88).

FROZEN = ("decoder.pos_emb.weight",)


def noam_factor(step: int, d_model: int, warmup: int) -> float:
    """base_trainer.py:114-117 (the scheduler's lambda; `step` counts completed optimizer steps)."""
    s = step + 1
    return (d_model ** -0.5) * min(s ** -0.5, s * warmup ** -1.5)


def xe_loss(w: Weights, model_cfg, vocab, feats: Tensor, tokens: Tensor, targets: Tensor,
            boxes: Optional[Tensor] = None) -> Tensor:
    """loss_fn(model(items).view(-1, V), shifted_right_caption_tokens.view(-1)), vi_trainer.py:107-111 (autograd on)."""
    enc, enc_mask = encode(w, model_cfg, feats, boxes)
    logp = decode(w, model_cfg, tokens, enc, enc_mask, vocab.padding_idx, None)
    return F.nll_loss(logp.reshape(-1, logp.shape[-1]), targets.reshape(-1), ignore_index=vocab.padding_idx)


def xe_train_steps(w: Weights, model_cfg, vocab, batches, lr: float, warmup: int, bf16_linear_weights: bool = False,
                   dropout_seeds=None):
    """Runs len(batches) optimizer steps in place on a float32 copy of `w`.  batches: (feats, tokens, targets[, boxes]).
    Returns (weights after the last step, [loss per step], gradients of the FIRST step).

    ``bf16_linear_weights`` is the mixed-precision yardstick for the GPU trainer: the forward and backward passes see
    every Linear weight rounded to bf16 (straight-through: the gradient is applied to the fp32 master weight), which is
    what a bf16 shadow copy of fp32 master weights does.  The synthetic start weights are bf16-exact, so the first step
    is unaffected; from the second step on the updated weights no longer are."""
    params = {k: v.detach().clone().float().requires_grad_(k not in FROZEN) for k, v in w.items()}
    trainable = [p for p in params.values() if p.requires_grad]
    optim = torch.optim.Adam(trainable, lr=lr, betas=(0.9, 0.98))
    sched = torch.optim.lr_scheduler.LambdaLR(optim, lambda s: noam_factor(s, model_cfg.ENCODER.D_MODEL, warmup))
    emb, pad = "decoder.word_emb.components.weight", vocab.padding_idx
    losses, first_grads = [], None
    for batch in batches:
        feats, tokens, targets = batch[:3]
        boxes = batch[3] if len(batch) > 3 else None
        optim.zero_grad()
        seen = params
        if bf16_linear_weights:
            seen = {k: (p + (p.detach().to(torch.bfloat16).float() - p.detach()))
                    if (p.dim() == 2 and k.endswith(".weight") and "_emb" not in k) else p for k, p in params.items()}
        with train_dropout(None if dropout_seeds is None else dropout_seeds[len(losses)]):
            loss = xe_loss(seen, model_cfg, vocab, feats, tokens, targets, boxes)
        loss.backward()
        if params[emb].grad is not None:
            params[emb].grad[pad].zero_()          # nn.Embedding(padding_idx=pad)
        if first_grads is None:
            first_grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in params.items()}
        optim.step()
        sched.step()
        losses.append(float(loss.detach()))
    return {k: v.detach() for k, v in params.items()}, losses, first_grads


# The self-critical step (trainers/vi_trainer.py:121-151): captions and their log-probs come from a beam search run with
# autograd on, the reward of each caption from CIDEr, and
#     loss = mean over (B, b) of  -(mean_T log_probs) * (reward - mean_b reward)                       (:146-148)
# The log-probs the reference differentiates are produced step by step by the stateful decoder; each equals the
# teacher-forced log-prob of the same token given the same prefix, and positions of finished beams hold <pad> with a
# constant 0 (beam_search.py:49-55).  This restatement therefore evaluates the FINAL sequences teacher-forced -- one
# decoder pass instead of max_len stateful steps -- and gen_golden_train.py checks loss, per-token log-probs and every
# gradient against the reference's own backward through its beam search (dropout p = 0).

def scst_loss(w: Weights, model_cfg, vocab, feats: Tensor, captions: Tensor, rewards: Tensor,
              boxes: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """captions (B, b, T) int64 as beam_search(out_size=b) returns them, rewards (B, b).  -> (loss, token log-probs (B, b, T))."""
    b_s, beam, t_len = captions.shape
    enc, enc_mask = encode(w, model_cfg, feats, boxes)
    enc = enc.repeat_interleave(beam, 0)
    enc_mask = enc_mask.repeat_interleave(beam, 0)
    seqs = captions.reshape(b_s * beam, t_len)
    tokens = torch.cat([torch.full((b_s * beam, 1), vocab.bos_idx, dtype=torch.long), seqs[:, :-1]], dim=1)
    logp = decode(w, model_cfg, tokens, enc, enc_mask, vocab.padding_idx, None)
    tok_lp = logp.gather(-1, seqs.unsqueeze(-1)).squeeze(-1) * (seqs != vocab.padding_idx)
    tok_lp = tok_lp.view(b_s, beam, t_len)
    loss = (-tok_lp.mean(-1) * (rewards - rewards.mean(-1, keepdim=True))).mean()
    return loss, tok_lp


def scst_step(w: Weights, model_cfg, vocab, feats: Tensor, captions: Tensor, rewards: Tensor, rl_lr: float,
              boxes: Optional[Tensor] = None, bf16_linear_weights: bool = False):
    """One update with Adam(lr=rl_lr), default betas, no scheduler (vi_trainer.py:213).  -> (weights, loss, grads)."""
    params = {k: v.detach().clone().float().requires_grad_(k not in FROZEN) for k, v in w.items()}
    optim = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=rl_lr)
    seen = params
    if bf16_linear_weights:
        seen = {k: (p + (p.detach().to(torch.bfloat16).float() - p.detach()))
                if (p.dim() == 2 and k.endswith(".weight") and "_emb" not in k) else p for k, p in params.items()}
    loss, _ = scst_loss(seen, model_cfg, vocab, feats, captions, rewards, boxes)
    loss.backward()
    emb = "decoder.word_emb.components.weight"
    if params[emb].grad is not None:
        params[emb].grad[vocab.padding_idx].zero_()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in params.items()}
    optim.step()
    return {k: v.detach() for k, v in params.items()}, float(loss.detach()), grads
