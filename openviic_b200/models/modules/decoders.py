"""Decoders (reference: models/modules/decoders.py:13-173), running on the sm_100a kernels.

These classes keep the reference's module-level semantics (stateful caches of raw inputs, the same
registered states) so they are drop-in at the registry level; the optimised whole-path decode
(K/V projected once, ancestry-indirected cache, fused log-softmax/top-k) lives in the engine that
``BaseTransformer.beam_search`` drives.
"""

from __future__ import annotations

import torch
from torch import nn

from ... import ops
from ...builders.decoder_builder import META_DECODER
from ...builders.text_embedding_builder import build_text_embedding
from ..utils import generate_padding_mask, generate_sequential_mask, sinusoid_encoding_table
from .attentions import MultiHeadAttention
from .containers import Module, ModuleList
from .positionwise_feed_forward import PositionWiseFeedForward


class DecoderLayer(Module):
    """Self-attention, cross-attention, feed-forward, zero <pad> rows (decoders.py:13-28)."""

    def __init__(self, config):
        super().__init__()
        self.self_attn = MultiHeadAttention(config.SELF_ATTENTION)
        self.enc_attn = MultiHeadAttention(config.ENC_ATTENTION)
        self.pwff = PositionWiseFeedForward(config.ENC_ATTENTION)

    def forward(self, queries, keys, values, self_padding_mask, self_attention_mask, enc_attention_mask, **kwargs):
        self_att = self.self_attn(queries, queries, queries, padding_mask=self_padding_mask,
                                  attention_mask=self_attention_mask, **kwargs)
        enc_att = self.enc_attn(self_att, keys, values, padding_mask=self_padding_mask,
                                attention_mask=enc_attention_mask, **kwargs)
        return self.pwff(enc_att, zero_rows=self_padding_mask.squeeze(1).squeeze(1))


class MeshedDecoderLayer(Module):
    """Cross-attends to every encoder level with SHARED enc_attn weights and mixes the results with
    sigmoid gates: sum_i sigmoid(W_i [s; c_i]) * c_i / sqrt(N) (decoders.py:30-73)."""

    def __init__(self, config):
        super().__init__()
        self.self_attn = MultiHeadAttention(config.SELF_ATTENTION)
        self.enc_attn = MultiHeadAttention(config.ENC_ATTENTION)
        self.pwff = PositionWiseFeedForward(config.ENC_ATTENTION)
        self.fc_alphas = nn.ModuleList([nn.Linear(2 * config.D_MODEL, config.D_MODEL)
                                        for _ in range(config.N_ENCODER_LAYERS)])
        self.nlayers = config.N_ENCODER_LAYERS
        for fc in self.fc_alphas:
            nn.init.xavier_uniform_(fc.weight)
            nn.init.constant_(fc.bias, 0)

    def forward(self, queries, keys, values, self_padding_mask, self_attention_mask, enc_attention_mask, **kwargs):
        with torch.no_grad():
            self_att = self.self_attn(queries, queries, queries, padding_mask=self_padding_mask,
                                      attention_mask=self_attention_mask, **kwargs)
            enc_atts, gates = [], []
            for ith in range(self.nlayers):
                level = keys[:, ith]
                enc_att = self.enc_attn(self_att, level, level, padding_mask=self_padding_mask,
                                        attention_mask=enc_attention_mask, **kwargs)
                fc = self.fc_alphas[ith]
                gates.append(ops.linear(torch.cat([self_att, enc_att], dim=-1), ops.cached_bf16(fc.weight), fc.bias,
                                        out_dtype=torch.float32))
                enc_atts.append(enc_att)
            mixed = ops.meshed_mix(torch.stack(gates, 0), torch.stack(enc_atts, 0))
            return self.pwff(mixed, zero_rows=self_padding_mask.squeeze(1).squeeze(1))


class _DecoderBase(Module):
    layer_cls = DecoderLayer

    def __init__(self, config, vocab):
        super().__init__()
        self.d_model = config.D_MODEL
        self.max_len = vocab.max_caption_length
        self.padding_idx = vocab.padding_idx
        self.N = config.LAYERS
        self.word_emb = build_text_embedding(config.TEXT_EMBEDDING, vocab)
        self.pos_emb = nn.Embedding.from_pretrained(
            sinusoid_encoding_table(max_len=self.max_len + 1, d_model=config.D_MODEL, padding_idx=0), freeze=True)
        self.layers = ModuleList([self.layer_cls(config.ATTENTION) for _ in range(config.LAYERS)])
        self.fc = nn.Linear(config.D_MODEL, len(vocab), bias=False)
        self.register_state("running_mask_self_attention", torch.zeros((1, 1, 0)).bool())
        self.register_state("running_seq", torch.zeros((1,)).long())

    def forward(self, caption_tokens, encoder_features, encoder_attention_mask):
        """tokens (R,S) int64 -> log-probs (R,S,V) fp32 (decoders.py:95-123 / :145-173)."""
        with torch.no_grad():
            b_s, seq_len = caption_tokens.shape[:2]
            padding_masks = generate_padding_mask(caption_tokens, self.padding_idx).to(caption_tokens.device)
            self_attention_masks = torch.logical_or(padding_masks,
                                                    generate_sequential_mask(seq_len).to(caption_tokens.device))
            if self._is_stateful:
                self.running_mask_self_attention = torch.cat(
                    [self.running_mask_self_attention, self_attention_masks], -1)
                self_attention_masks = self.running_mask_self_attention
            seq = torch.arange(1, seq_len + 1, device=caption_tokens.device).view(1, -1).expand(b_s, -1)
            seq = seq.masked_fill(padding_masks.squeeze(1).squeeze(1), 0)
            if self._is_stateful:
                self.running_seq.add_(1)  # not zeroed for <pad> rows, exactly like the reference
                seq = self.running_seq
            embedded, _ = self.word_emb(caption_tokens)
            out = (embedded + self.pos_emb(seq)).float()
            for layer in self.layers:
                out = layer(queries=out, keys=encoder_features, values=encoder_features,
                            self_padding_mask=padding_masks, self_attention_mask=self_attention_masks,
                            enc_attention_mask=encoder_attention_mask)
            logits = ops.linear(out, ops.cached_bf16(self.fc.weight), None, out_dtype=torch.float32)
            return ops.log_softmax(logits)


@META_DECODER.register()
class Decoder(_DecoderBase):
    layer_cls = DecoderLayer


@META_DECODER.register()
class MeshedDecoder(_DecoderBase):
    layer_cls = MeshedDecoderLayer
