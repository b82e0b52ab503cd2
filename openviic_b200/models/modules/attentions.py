"""Attention variants + the multi-head wrapper, running on the sm_100a kernels.

Drop-in for the reference's models/modules/attentions.py: same registry names, constructor
``(config)``, forward signatures, parameter names and initialisers, so a reference ``state_dict``
loads unchanged.  The arithmetic is: projections on the tcgen05 GEMM (``ops.linear``), the
softmax(QK^T)V core in the fused attention kernel (``ops.attention``), residual+LayerNorm and the
AoA gate in the row kernels.  Inference only (dropout is the identity, no autograd).
"""

from __future__ import annotations

import math

import torch
from torch import nn

from ... import ops
from ...builders.attention_builder import META_ATTENTION, build_attention
from .containers import Module


class _ProjectedAttention(nn.Module):
    """fc_q / fc_k / fc_v / fc_o shared by every variant (attentions.py:14-42)."""

    def __init__(self, config):
        super().__init__()
        self.d_model, self.h = config.D_MODEL, config.HEAD
        self.d_k, self.d_v = config.D_KEY, config.D_VALUE
        if self.d_k != 64 or self.d_v != 64:
            raise ValueError("the sm_100a attention kernels are specialised for D_KEY = D_VALUE = 64")
        self.fc_q = nn.Linear(self.d_model, self.h * self.d_k)
        self.fc_k = nn.Linear(self.d_model, self.h * self.d_k)
        self.fc_v = nn.Linear(self.d_model, self.h * self.d_v)
        self.fc_o = nn.Linear(self.h * self.d_v, self.d_model)
        self.init_weights()

    def init_weights(self):
        for fc in (self.fc_q, self.fc_k, self.fc_v, self.fc_o):
            nn.init.xavier_uniform_(fc.weight)
            nn.init.constant_(fc.bias, 0)

    def _project(self, queries, keys, values):
        hd = self.h * self.d_k
        if queries is keys and keys is values:  # self-attention: one stacked q|k|v GEMM
            w = ops.cached_bf16(self.fc_q.weight, self.fc_k.weight, self.fc_v.weight)
            b = ops.cached_f32_cat(self.fc_q.bias, self.fc_k.bias, self.fc_v.bias)
            qkv = ops.linear(queries, w, b)
            return qkv[..., :hd], qkv[..., hd:2 * hd], qkv[..., 2 * hd:]
        q = ops.linear(queries, ops.cached_bf16(self.fc_q.weight), self.fc_q.bias)
        if keys is values:
            w = ops.cached_bf16(self.fc_k.weight, self.fc_v.weight)
            b = ops.cached_f32_cat(self.fc_k.bias, self.fc_v.bias)
            kv = ops.linear(keys, w, b)
            return q, kv[..., :hd], kv[..., hd:]
        k = ops.linear(keys, ops.cached_bf16(self.fc_k.weight), self.fc_k.bias)
        v = ops.linear(values, ops.cached_bf16(self.fc_v.weight), self.fc_v.bias)
        return q, k, v

    def _attend(self, queries, keys, values, attention_mask, **core):
        with torch.no_grad():
            q, k, v = self._project(queries, keys, values)
            out = ops.attention(q, k, v, self.h, mask=attention_mask, scale=1.0 / math.sqrt(self.d_k), **core)
            return ops.linear(out, ops.cached_bf16(self.fc_o.weight), self.fc_o.bias, out_dtype=torch.float32)


@META_ATTENTION.register()
class ScaledDotProductAttention(_ProjectedAttention):
    """attentions.py:9-58."""

    def forward(self, queries, keys, values, attention_mask=None):
        return self._attend(queries, keys, values, attention_mask)


@META_ATTENTION.register()
class AugmentedGeometryScaledDotProductAttention(_ProjectedAttention):
    """Box-relation biased attention, attentions.py:61-114: logits += log(clamp(g, 1e-6))."""

    def forward(self, queries, keys, values, relative_geometry_weights, attention_mask=None):
        return self._attend(queries, keys, values, attention_mask, geometry=relative_geometry_weights)


@META_ATTENTION.register()
class AugmentedMemoryScaledDotProductAttention(_ProjectedAttention):
    """Memory-augmented attention, attentions.py:117-185: m learned K/V slots appended, never masked."""

    def __init__(self, config):
        self.m = config.MEMORY
        super().__init__(config)

    def init_weights(self):
        if not hasattr(self, "m_k"):
            self.m_k = nn.Parameter(torch.empty(1, self.m, self.h * self.d_k))
            self.m_v = nn.Parameter(torch.empty(1, self.m, self.h * self.d_v))
        super().init_weights()
        nn.init.normal_(self.m_k, 0, 1 / self.d_k)
        nn.init.normal_(self.m_v, 0, 1 / self.m)

    def forward(self, queries, keys, values, attention_mask=None):
        mem_k = (math.sqrt(self.d_k) * self.m_k.detach()).to(torch.bfloat16)
        mem_v = (math.sqrt(self.m) * self.m_v.detach()).to(torch.bfloat16)
        return self._attend(queries, keys, values, attention_mask, mem_k=mem_k, mem_v=mem_v)


@META_ATTENTION.register()
class AdaptiveScaledDotProductAttention(_ProjectedAttention):
    """Adaptive attention (attentions.py:188-268): every query also attends to its own projected language signal --
    one extra softmax column q_i . s_i / sqrt(d_k) whose value is s_i (the reference builds it with Python loops over
    the queries, :255-263).  Only AdaptiveDecoder uses it and that class cannot be constructed in the reference
    (SURVEY.md section 8c), so parity is pinned at the operator level: tests/golden/adaptive_attention.npz holds the
    reference class's own output."""

    def __init__(self, config):
        super().__init__(config)
        self.fc_s = nn.Linear(self.d_model, self.h * self.d_k)
        nn.init.xavier_uniform_(self.fc_s.weight)
        nn.init.constant_(self.fc_s.bias, 0)

    def forward(self, queries, keys, values, language_signals, attention_mask=None):
        with torch.no_grad():
            signals = ops.linear(language_signals, ops.cached_bf16(self.fc_s.weight), self.fc_s.bias)
        return self._attend(queries, keys, values, attention_mask, sentinel=signals)


class MultiHeadAttention(Module):
    """Attention + residual LayerNorm (+ attention-on-attention gate), attentions.py:270-317."""

    def __init__(self, config):
        super().__init__()
        d_model = config.D_MODEL
        self.use_aoa = config.USE_AOA
        if self.use_aoa:
            self.informative_attention = nn.Linear(2 * d_model, d_model)
            self.gated_attention = nn.Linear(2 * d_model, d_model)
        self.attention = build_attention(config)
        self.dropout = nn.Dropout(p=config.DROPOUT)
        self.layer_norm = nn.LayerNorm(d_model)
        self.can_be_stateful = config.CAN_BE_STATEFUL
        if self.can_be_stateful:
            self.register_state("running_keys", torch.zeros((0, d_model)))
            self.register_state("running_values", torch.zeros((0, d_model)))

    def forward(self, queries, keys, values, padding_mask, attention_mask, **kwargs):
        with torch.no_grad():
            self_attention = keys is queries and values is queries
            # activations stay fp32 between modules (residual stream); GEMMs round their operands to bf16
            if self.can_be_stateful and self._is_stateful:
                # the reference's cache semantics: raw inputs appended along time (attentions.py:297-302)
                self.running_keys = torch.cat([self.running_keys.to(keys.dtype), keys], 1)
                self.running_values = self.running_keys if self_attention else torch.cat(
                    [self.running_values.to(values.dtype), values], 1)
                keys, values = self.running_keys, self.running_values
            out = self.attention(queries, keys, values, attention_mask=attention_mask, **kwargs)
            out = ops.add_layernorm(out, queries, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
            if self.use_aoa:
                w = ops.cached_bf16(self.informative_attention.weight, self.gated_attention.weight)
                b = ops.cached_f32_cat(self.informative_attention.bias, self.gated_attention.bias)
                gates = ops.linear(torch.cat([queries, out], dim=-1), w, b, out_dtype=torch.float32)
                out = ops.aoa_gate(gates)
            return out
