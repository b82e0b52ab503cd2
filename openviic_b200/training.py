"""T1 -- the cross-entropy training step on the sm_100a kernels.

Reference: the body of ``Trainer.train`` (trainers/vi_trainer.py:105-119) -- ``out = model(items)``,
``NLLLoss(ignore_index=<pad>)`` against the shifted tokens, ``backward``, ``Adam(lr, betas=(0.9, 0.98)).step()``,
``LambdaLR(lambda_lr).step()`` (trainers/base_trainer.py:89-91, 114-117) -- for the standard transformer: FeatureEmbedding
-> Encoder (ScaledDotProductAttention) -> Decoder (models/standard_stransformer.py:21-31).

Everything numeric runs in hand-written kernels through the C ABI: the forward pass on ``cap_linear`` (tcgen05 GEMM),
``cap_attention``, ``cap_train_layernorm_fwd``, ``cap_train_embed_fwd``; the backward pass on ``cap_linear`` again (dX =
dY.W and dW = dY^T.X over operands transposed by ``cap_transpose_bf16``, which also yields the bias gradients),
``cap_attention_backward``, ``cap_train_layernorm_bwd``, ``cap_train_relu_bwd``, ``cap_train_embed_bwd``; the loss and
its gradient in ``cap_train_xent``; the update in ``cap_train_adam``.  torch supplies device memory and the stream.

Precision: fp32 master weights, Adam moments, gradients, residual stream, LayerNorm and loss; bf16 GEMM / attention
operands (a bf16 shadow copy of the weights is rewritten by the optimizer kernel) -- the mixed-precision recipe.

Dropout: the reference's nn.Dropout modules (vision_embeddings.py:18, attentions.py:308,
positionwise_feed_forward.py:24-25) draw from torch's generator.  Here the same sites, probabilities and scaling run on a
counter-based mask (``cap_train_dropout``: element i of the tensor at the module named ``site`` is kept iff
hash(i, seed, crc32(site)) >= p * 2^32), regenerated in the backward pass instead of stored; ``dropout_seed`` + the step
number seeds each step.  The oracle restates the hash, and the fixture generator pins it against the real reference with
its Dropout modules' forward wrapped to use the same mask.  Without a seed ``XETrainer`` refuses configs with dropout
unless ``ignore_dropout=True`` (the p = 0 computation).
"""

from __future__ import annotations

import ctypes as C
import math
import zlib
from typing import Dict, List, Optional, Tuple

import torch

from . import cabi
from .cabi import ACT_NONE, ACT_RELU, CAP_BF16, CAP_F32
from .models.utils import visual_position_table

Tensor = torch.Tensor
FROZEN = ("decoder.pos_emb.weight",)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def noam_factor(step: int, d_model: int, warmup: int) -> float:
    """base_trainer.py:114-117; ``step`` = optimizer steps already taken."""
    s = step + 1
    return (d_model ** -0.5) * min(s ** -0.5, s * warmup ** -1.5)


class _Linear:
    """One (possibly stacked) Linear: views into the flat buffers."""
    __slots__ = ("w16", "b32", "gw", "gb", "wt16", "n", "k")

    def __init__(self, w16, b32, gw, gb):
        self.w16, self.b32, self.gw, self.gb = w16, b32, gw, gb
        self.n, self.k = w16.shape
        self.wt16 = torch.zeros((self.k, (self.n + 7) // 8 * 8), device=w16.device, dtype=torch.bfloat16)


class XETrainer:
    def __init__(self, model, lr: float = 1.0, warmup: int = 10000, betas: Tuple[float, float] = (0.9, 0.98),
                 eps: float = 1e-8, ignore_dropout: bool = False, dropout_seed: Optional[int] = None):
        cfg = model.model_config
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("XETrainer runs on a CUDA device only (no CPU fallback)")
        enc, dec = cfg.ENCODER, cfg.DECODER
        att_cfgs = (enc.SELF_ATTENTION, dec.ATTENTION.SELF_ATTENTION, dec.ATTENTION.ENC_ATTENTION)
        supported = (cfg.VISION_EMBEDDING.ARCHITECTURE == "FeatureEmbedding" and enc.ARCHITECTURE == "Encoder"
                     and dec.ARCHITECTURE == "Decoder" and dec.TEXT_EMBEDDING.ARCHITECTURE == "UsualEmbedding"
                     and dec.TEXT_EMBEDDING.WORD_EMBEDDING is None
                     and all(a.ARCHITECTURE == "ScaledDotProductAttention" and not a.USE_AOA for a in att_cfgs))
        if not supported:
            raise NotImplementedError("XETrainer covers the standard transformer (FeatureEmbedding, Encoder, Decoder, plain "
                                      "scaled dot-product attention); other registry variants have no backward kernels")
        drops = [cfg.VISION_EMBEDDING.DROPOUT] + [a.DROPOUT for a in att_cfgs]
        if any(float(p) != 0.0 for p in drops) and dropout_seed is None and not ignore_dropout:
            raise NotImplementedError("this config trains with dropout: pass dropout_seed=<int> (counter-based masks), or "
                                      "ignore_dropout=True to run it as if DROPOUT were 0")
        self.dropout_seed = None if ignore_dropout else dropout_seed
        self.p_vision, self.p_enc, self.p_self, self.p_cross = (float(p) for p in drops)
        self.model, self.vocab = model, model.vocab
        self.d, self.heads, self.dff = enc.D_MODEL, enc.SELF_ATTENTION.HEAD, enc.SELF_ATTENTION.D_FF
        self.enc_layers, self.dec_layers = enc.LAYERS, dec.LAYERS
        self.V, self.pad = len(model.vocab), model.vocab.padding_idx
        self.ldv = (self.V + 7) // 8 * 8
        self.lr, self.warmup, self.betas, self.eps = float(lr), int(warmup), betas, float(eps)
        self.steps_done = 0
        self._build_flat(model.state_dict())
        self.pos_words = model.state_dict()["decoder.pos_emb.weight"].detach().float().contiguous()
        self._visual_pos: Dict[int, Tensor] = {}

    # ------------------------------------------------------------------------------------------------ parameters
    def _build_flat(self, state: Dict[str, Tensor]) -> None:
        names = [k for k, v in state.items() if v.dtype.is_floating_point and v.numel() and k not in FROZEN
                 and "running_" not in k]
        order: List[str] = []
        for k in names:   # q|k|v (and their biases) adjacent, so that the stacked projections are plain views
            if k.endswith("fc_q.weight"):
                p = k[:-len("fc_q.weight")]
                order += [p + f"fc_{x}.weight" for x in "qkv"] + [p + f"fc_{x}.bias" for x in "qkv"]
            elif any(k.endswith(f"fc_{x}.{y}") for x in "qkv" for y in ("weight", "bias")):
                continue
            else:
                order.append(k)
        assert sorted(order) == sorted(names)
        offsets, total = {}, 0
        for k in order:
            assert state[k].numel() % 64 == 0 or not k.endswith(("fc_q.weight", "fc_k.weight", "fc_q.bias", "fc_k.bias")), k
            offsets[k] = total
            total += state[k].numel()
            if not k.endswith(("fc_q.weight", "fc_k.weight", "fc_q.bias", "fc_k.bias")):
                total = (total + 63) // 64 * 64
        dev = self.device
        self.p32 = torch.zeros(total, device=dev, dtype=torch.float32)
        self.g32 = torch.zeros(total, device=dev, dtype=torch.float32)
        self.m32 = torch.zeros(total, device=dev, dtype=torch.float32)
        self.v32 = torch.zeros(total, device=dev, dtype=torch.float32)
        self.p16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        self.offsets, self.shapes = offsets, {k: tuple(state[k].shape) for k in order}
        for k in order:
            self.p32[offsets[k]:offsets[k] + state[k].numel()].copy_(state[k].detach().reshape(-1).float())
        self.p16.copy_(self.p32)

    def _view(self, flat: Tensor, name: str, rows: Optional[int] = None) -> Tensor:
        """View of one parameter; ``rows`` stacks the following parameters of the same width (q|k|v)."""
        shape = self.shapes[name]
        n = shape[0] if rows is None else rows
        tail = shape[1:]
        count = n * (int(math.prod(tail)) if tail else 1)
        return flat[self.offsets[name]:self.offsets[name] + count].view((n,) + tuple(tail))

    def _linear(self, prefix: str, first: str = "", stack: int = 1, bias: bool = True) -> _Linear:
        name = prefix + first
        rows = self.shapes[name + ".weight"][0] * stack
        return _Linear(self._view(self.p16, name + ".weight", rows), self._view(self.p32, name + ".bias", rows) if bias else None,
                       self._view(self.g32, name + ".weight", rows), self._view(self.g32, name + ".bias", rows) if bias else None)

    def _linears(self):
        """(proj, encoder layers, decoder layers, vocabulary projection) as _Linear views; built on first use and kept
        (the flat buffers never move, and each _Linear owns the scratch for its transposed weight)."""
        if getattr(self, "_lin_cache", None) is None:
            L = self.enc_layers
            proj = self._linear("vision_embedding.proj")
            enc = [dict(qkv=self._linear(f"encoder.layers.{l}.mhatt.attention.", "fc_q", 3),
                        o=self._linear(f"encoder.layers.{l}.mhatt.attention.fc_o"),
                        fc1=self._linear(f"encoder.layers.{l}.pwff.fc1"), fc2=self._linear(f"encoder.layers.{l}.pwff.fc2")) for l in range(L)]
            dec = [dict(qkv=self._linear(f"decoder.layers.{l}.self_attn.attention.", "fc_q", 3),
                        o1=self._linear(f"decoder.layers.{l}.self_attn.attention.fc_o"),
                        q=self._linear(f"decoder.layers.{l}.enc_attn.attention.fc_q"),
                        kv=self._linear(f"decoder.layers.{l}.enc_attn.attention.", "fc_k", 2),
                        o2=self._linear(f"decoder.layers.{l}.enc_attn.attention.fc_o"),
                        fc1=self._linear(f"decoder.layers.{l}.pwff.fc1"), fc2=self._linear(f"decoder.layers.{l}.pwff.fc2"))
                   for l in range(self.dec_layers)]
            fc = self._linear("decoder.fc", bias=False)
            self._lin_cache = (proj, enc, dec, fc)
        return self._lin_cache

    def parameters(self) -> Dict[str, Tensor]:
        """fp32 master weights by state_dict name (views)."""
        return {k: self._view(self.p32, k) for k in self.shapes}

    def gradients(self) -> Dict[str, Tensor]:
        return {k: self._view(self.g32, k) for k in self.shapes}

    def sync_to_model(self) -> None:
        """Copies the master weights back into the model's parameters (e.g. before ``model.beam_search``)."""
        with torch.no_grad():
            state = self.model.state_dict()
            for k, v in self.parameters().items():
                state[k].copy_(v)

    # ------------------------------------------------------------------------------------------------ kernels
    def _gemm(self, x16: Tensor, w16: Tensor, bias: Optional[Tensor], out_f32: bool, act: int = ACT_NONE,
              k: Optional[int] = None) -> Tensor:
        """act(x . w^T + bias): x (M, K) bf16 with row stride ldx, w (N, K) bf16 dense -> (M, ceil8(N))."""
        m = x16.shape[0]
        n = w16.shape[0]
        kk = w16.shape[1] if k is None else k
        ldy = (n + 7) // 8 * 8
        y = torch.empty((m, ldy), device=self.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
        if ldy != n:
            y.zero_()
        cabi.call("cap_linear", x16.data_ptr(), x16.stride(0), w16.data_ptr(), None if bias is None else bias.data_ptr(),
                  y.data_ptr(), ldy, CAP_F32 if out_f32 else CAP_BF16, act, m, n, kk, _stream())
        return y

    def _transpose(self, x16: Tensor, rows: int, cols: int, colsum: Optional[Tensor] = None) -> Tensor:
        """(rows, cols) bf16 with row stride -> (cols, ceil8(rows)), zero padded; colsum += column sums."""
        ldo = (rows + 7) // 8 * 8
        out = torch.empty((cols, ldo), device=self.device, dtype=torch.bfloat16)
        cabi.call("cap_transpose_bf16", x16.data_ptr(), x16.stride(0), out.data_ptr(), ldo,
                  None if colsum is None else colsum.data_ptr(), rows, cols, _stream())
        return out

    def _lin_fwd(self, lin: _Linear, x16: Tensor, out_f32: bool, act: int = ACT_NONE) -> Tensor:
        return self._gemm(x16, lin.w16, lin.b32, out_f32, act)

    def _lin_bwd(self, lin: _Linear, x16: Tensor, dy16: Tensor, need_dx: bool = True, dx_f32: bool = True) -> Optional[Tensor]:
        """dW = dY^T . X and db = colsum(dY) into the gradient buffer (every Linear is used once per step, so the GEMM
        writes the gradient in place); returns dX = dY . W.  dy16 (M, >= N) bf16, columns beyond N zero."""
        m, n, k = x16.shape[0], lin.n, lin.k
        dyt = self._transpose(dy16, m, n, lin.gb)                       # (N, M8)
        xt = self._transpose(x16, m, k)                                  # (K, M8)
        m8 = dyt.shape[1]
        # few output tiles over a long contraction: split the M axis over enough CTAs to fill the GPU
        tiles = ((n + 127) // 128) * ((k + 255) // 256 if k >= 256 else 1)
        kb = (m8 + 63) // 64
        splits = max(1, min(16, kb, -(-148 // tiles)))
        per = -(-kb // splits)
        splits = -(-kb // per)
        if splits > 1:
            parts = torch.empty((splits, n, k), device=self.device, dtype=torch.float32)
            cabi.call("cap_linear_splitk", dyt.data_ptr(), m8, xt.data_ptr(), parts.data_ptr(), k, n, k, m8, splits, _stream())
            cabi.call("cap_sum_partials", parts.data_ptr(), splits, n * k, lin.gw.data_ptr(), _stream())
        else:
            cabi.call("cap_linear", dyt.data_ptr(), m8, xt.data_ptr(), None, lin.gw.data_ptr(), k, CAP_F32, ACT_NONE, n, k, m8, _stream())
        if not need_dx:
            return None
        n8 = lin.wt16.shape[1]
        dy_view = dy16 if dy16.shape[1] == n8 else dy16[:, :n8]
        return self._gemm(dy_view, lin.wt16, None, dx_f32, k=n8)

    def _refresh_transposed(self, lins: List[_Linear]) -> None:
        for lin in lins:   # W^T (K, ceil8(N)) for dX = dY . W; the pad columns stay zero
            cabi.call("cap_transpose_bf16", lin.w16.data_ptr(), lin.k, lin.wt16.data_ptr(), lin.wt16.shape[1], None, lin.n, lin.k,
                      _stream())

    def _ln_fwd(self, a32, res32, prefix, pos=None, zero_rows=None):
        rows = a32.shape[0]
        pre = torch.empty((rows, self.d), device=self.device, dtype=torch.float32)
        out32 = torch.empty_like(pre)
        out16 = torch.empty((rows, self.d), device=self.device, dtype=torch.bfloat16)
        g, b = self._view(self.p32, prefix + ".weight"), self._view(self.p32, prefix + ".bias")
        cabi.call("cap_train_layernorm_fwd", a32.data_ptr(), None if res32 is None else res32.data_ptr(), g.data_ptr(), b.data_ptr(),
                  1e-5, None if pos is None else pos.data_ptr(), 0 if pos is None else pos.shape[0],
                  None if zero_rows is None else zero_rows.data_ptr(), pre.data_ptr(), out32.data_ptr(), out16.data_ptr(), rows,
                  self.d, _stream())
        return pre, out32, out16

    def _ln_bwd(self, dout_a, dout_b, pre, prefix, zero_rows=None):
        rows = pre.shape[0]
        d32 = torch.empty_like(pre)
        d16 = torch.empty((rows, self.d), device=self.device, dtype=torch.bfloat16)
        g = self._view(self.p32, prefix + ".weight")
        cabi.call("cap_train_layernorm_bwd", dout_a.data_ptr(), None if dout_b is None else dout_b.data_ptr(), pre.data_ptr(),
                  g.data_ptr(), 1e-5, None if zero_rows is None else zero_rows.data_ptr(), d32.data_ptr(), d16.data_ptr(),
                  self._view(self.g32, prefix + ".weight").data_ptr(), self._view(self.g32, prefix + ".bias").data_ptr(), rows, self.d,
                  _stream())
        return d32, d16

    def _dropout(self, t: Tensor, site: str, p: float) -> None:
        """In place; forward and backward call it with the same site (the mask is regenerated, not stored)."""
        if self.dropout_seed is None or p == 0.0:
            return
        assert t.is_contiguous()
        seed = (self.dropout_seed + self.steps_done + getattr(self, "rl_steps_done", 0)) & 0xFFFFFFFF
        cabi.call("cap_train_dropout", t.data_ptr(), CAP_F32 if t.dtype == torch.float32 else CAP_BF16, t.numel(),
                  int(p * 4294967296.0), 1.0 / (1.0 - p), seed, zlib.crc32(site.encode()), _stream())

    def _att_args(self, q, k, v, out, b, nq, nk, mask, mask_qs):
        return cabi.AttentionArgs(
            q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(),
            q_bs=nq * q.stride(0), k_bs=nk * k.stride(0), v_bs=nk * v.stride(0), o_bs=nq * out.stride(0),
            ldq=q.stride(0), ldk=k.stride(0), ldv=v.stride(0), ldo=out.stride(0),
            mask=mask.data_ptr(), mask_bs=mask.shape[1] * mask.shape[2], mask_qs=mask_qs, geometry=None, mem_k=None, mem_v=None,
            n_mem=0, B=b, H=self.heads, nq=nq, nk=nk, scale=1.0 / math.sqrt(64), sentinel=None, s_bs=0, lds=0)

    def _att_fwd(self, q, k, v, b, nq, nk, mask, mask_qs):
        """q (b*nq, hd) / k, v (b*nk, hd) row-strided bf16 views; mask uint8 (b, nq|1, nk)."""
        out = torch.empty((b * nq, self.heads * 64), device=self.device, dtype=torch.bfloat16)
        args = self._att_args(q, k, v, out, b, nq, nk, mask, mask_qs)
        cabi.call("cap_attention", C.byref(args), _stream())
        return out

    def _att_bwd(self, q, k, v, d_out, dq, dk, dv, b, nq, nk, mask, mask_qs):
        args = self._att_args(q, k, v, d_out, b, nq, nk, mask, mask_qs)   # `out` carries d_out's strides
        cabi.call("cap_attention_backward", C.byref(args), d_out.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), _stream())

    # ------------------------------------------------------------------------------------------------ one step
    def loss_and_grads(self, feats: Tensor, tokens: Tensor, targets: Tensor, row_weights: Optional[Tensor] = None) -> Tensor:
        """Forward + backward; gradients land in ``self.gradients()``; returns the loss (0-d fp32 device tensor).

        tokens / targets (B, T): one caption per image (the XE step).  (B, S, T): S captions per image that share the
        image's encoder output -- the S sequences of an image are S * T queries of one cross-attention problem; used by
        the self-critical step with ``row_weights`` (B * S,) fp32: loss = sum over sequences of weight * sum_t nll_t."""
        if not (feats.is_cuda and tokens.is_cuda and targets.is_cuda):
            raise RuntimeError("XETrainer.step takes CUDA tensors (no CPU fallback)")
        B, n, dfeat = feats.shape
        S = tokens.shape[1] if tokens.dim() == 3 else 1     # sequences per image
        T = tokens.shape[-1]
        want = (B, S, T) if tokens.dim() == 3 else (B, T)
        if tuple(tokens.shape) != want or tuple(targets.shape) != want or tokens.dtype != torch.int64 or targets.dtype != torch.int64:
            raise ValueError("tokens / targets must be int64 (B, T) or (B, S, T)")
        if n > 128 or S * T > 128 or T + 1 > self.pos_words.shape[0]:
            raise ValueError("at most 128 visual tokens, 128 caption tokens per image, max_caption_length tokens per caption")
        if row_weights is not None and (tuple(row_weights.shape) != (B * S,) or row_weights.dtype != torch.float32 or not row_weights.is_cuda):
            raise ValueError("row_weights must be a CUDA fp32 tensor (B * S,)")
        d, hd, L = self.d, self.heads * 64, self.enc_layers
        dev = self.device
        R = B * S                   # decoder sequences
        Me, Md = B * n, R * T
        self.g32.zero_()
        tokens = tokens.contiguous().view(-1)
        targets = targets.contiguous().view(-1)

        # the Linears (views into the flat buffers, built once); W^T refreshed from the current weights
        proj, enc, dec, fc = self._linears()
        self._refresh_transposed([x for layer in enc for x in layer.values()] + [x for layer in dec for x in layer.values()] + [fc])

        # ---------------- forward: encoder (vision_embeddings.py:15-20, encoders.py:17-40)
        f16 = torch.empty((Me, dfeat), device=dev, dtype=torch.bfloat16)
        enc_mask = torch.empty((B, 1, n), device=dev, dtype=torch.uint8)   # 1 = padded visual token
        src = feats.contiguous()
        cabi.call("cap_feature_mask_cast", src.data_ptr(), 1 if src.dtype == torch.float32 else 0, f16.data_ptr(), enc_mask.data_ptr(),
                  Me, dfeat, _stream())
        enc_rows = enc_mask.view(-1)
        if n not in self._visual_pos:
            self._visual_pos[n] = visual_position_table(n, d).to(dev).float().contiguous()
        x0 = self._lin_fwd(proj, f16, True)
        self._dropout(x0, "vision_embedding.dropout", self.p_vision)
        pre0, x32, x16 = self._ln_fwd(x0, None, "encoder.layer_norm", pos=self._visual_pos[n])
        enc_saved = []
        for l, w in enumerate(enc):
            p = f"encoder.layers.{l}."
            qkv = self._lin_fwd(w["qkv"], x16, False)
            att = self._att_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], B, n, n, enc_mask, 0)
            o32 = self._lin_fwd(w["o"], att, True)
            self._dropout(o32, p + "mhatt.dropout", self.p_enc)
            pre1, y32, y16 = self._ln_fwd(o32, x32, p + "mhatt.layer_norm")
            h16 = self._lin_fwd(w["fc1"], y16, False, ACT_RELU)
            self._dropout(h16, p + "pwff.dropout_2", self.p_enc)
            f32 = self._lin_fwd(w["fc2"], h16, True)
            self._dropout(f32, p + "pwff.dropout", self.p_enc)
            pre2, nx32, nx16 = self._ln_fwd(f32, y32, p + "pwff.layer_norm", zero_rows=enc_rows)
            enc_saved.append((x16, qkv, att, pre1, y16, h16, pre2))
            x32, x16 = nx32, nx16
        enc16 = x16

        # ---------------- forward: decoder (decoders.py:21-28, 95-123)
        pad_rows = (tokens == self.pad).to(torch.uint8)
        causal = torch.triu(torch.ones((T, T), device=dev, dtype=torch.bool), diagonal=1)
        self_mask = (causal.unsqueeze(0) | (tokens.view(R, 1, T) == self.pad)).to(torch.uint8).contiguous()   # (R, T, T)
        emb = self._view(self.p32, "decoder.word_emb.components.weight")
        e32 = torch.empty((Md, d), device=dev, dtype=torch.float32)
        e16 = torch.empty((Md, d), device=dev, dtype=torch.bfloat16)
        cabi.call("cap_train_embed_fwd", tokens.data_ptr(), emb.data_ptr(), self.pos_words.data_ptr(), T, self.pad, e32.data_ptr(),
                  e16.data_ptr(), Md, d, _stream())
        dec_saved = []
        for l, w in enumerate(dec):
            p = f"decoder.layers.{l}."
            qkv = self._lin_fwd(w["qkv"], e16, False)
            att1 = self._att_fwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], R, T, T, self_mask, T)
            o32 = self._lin_fwd(w["o1"], att1, True)
            self._dropout(o32, p + "self_attn.dropout", self.p_self)
            pre1, s32, s16 = self._ln_fwd(o32, e32, p + "self_attn.layer_norm")
            q16 = self._lin_fwd(w["q"], s16, False)
            kv16 = self._lin_fwd(w["kv"], enc16, False)
            att2 = self._att_fwd(q16, kv16[:, :hd], kv16[:, hd:], B, S * T, n, enc_mask, 0)   # an image's S sequences: S * T queries
            o32 = self._lin_fwd(w["o2"], att2, True)
            self._dropout(o32, p + "enc_attn.dropout", self.p_cross)
            pre2, c32, c16 = self._ln_fwd(o32, s32, p + "enc_attn.layer_norm")
            h16 = self._lin_fwd(w["fc1"], c16, False, ACT_RELU)
            self._dropout(h16, p + "pwff.dropout_2", self.p_cross)      # PositionWiseFeedForward(config.ENC_ATTENTION)
            f32 = self._lin_fwd(w["fc2"], h16, True)
            self._dropout(f32, p + "pwff.dropout", self.p_cross)
            pre3, ne32, ne16 = self._ln_fwd(f32, c32, p + "pwff.layer_norm", zero_rows=pad_rows)
            dec_saved.append((e16, qkv, att1, pre1, s16, q16, kv16, att2, pre2, c16, h16, pre3))
            e32, e16 = ne32, ne16
        logits = self._lin_fwd(fc, e16, True)                                  # (Md, ldv) fp32

        # ---------------- loss (base_trainer.py:91, vi_trainer.py:110)
        stats = torch.empty(2, device=dev, dtype=torch.float32)
        dlogits = torch.empty((Md, self.ldv), device=dev, dtype=torch.bfloat16)
        per_token = None if row_weights is None else row_weights.repeat_interleave(T).contiguous()
        cabi.call("cap_train_xent", logits.data_ptr(), self.ldv, targets.data_ptr(), self.pad, None if per_token is None else per_token.data_ptr(),
                  stats.data_ptr(), dlogits.data_ptr(), self.ldv, Md, self.V, _stream())
        loss = stats[1] / stats[0] if row_weights is None else stats[1].clone()

        # ---------------- backward: decoder
        g_a = self._lin_bwd(fc, e16, dlogits)                                  # (Md, d) fp32: gradient of the last layer's output
        g_b = None
        denc = torch.zeros((Me, d), device=dev, dtype=torch.float32)
        for l in reversed(range(self.dec_layers)):
            w, p = dec[l], f"decoder.layers.{l}."
            e16_in, qkv, att1, pre1, s16, q16, kv16, att2, pre2, c16, h16, pre3 = dec_saved[l]
            d3_32, d3_16 = self._ln_bwd(g_a, g_b, pre3, p + "pwff.layer_norm", zero_rows=pad_rows)
            self._dropout(d3_16, p + "pwff.dropout", self.p_cross)      # only the GEMM branch is dropped, not the residual
            dh16 = self._lin_bwd(w["fc2"], h16, d3_16, dx_f32=False)
            self._dropout(dh16, p + "pwff.dropout_2", self.p_cross)
            cabi.call("cap_train_relu_bwd", dh16.data_ptr(), h16.data_ptr(), dh16.numel(), _stream())
            dc_ffn = self._lin_bwd(w["fc1"], c16, dh16)
            d2_32, d2_16 = self._ln_bwd(d3_32, dc_ffn, pre2, p + "enc_attn.layer_norm")
            self._dropout(d2_16, p + "enc_attn.dropout", self.p_cross)
            datt2 = self._lin_bwd(w["o2"], att2, d2_16, dx_f32=False)
            dq16 = torch.empty_like(q16)
            dkv16 = torch.empty_like(kv16)
            self._att_bwd(q16, kv16[:, :hd], kv16[:, hd:], datt2, dq16, dkv16[:, :hd], dkv16[:, hd:], B, S * T, n, enc_mask, 0)
            ds_q = self._lin_bwd(w["q"], s16, dq16)
            denc_l = self._lin_bwd(w["kv"], enc16, dkv16)
            cabi.call("cap_axpy_f32", denc.data_ptr(), denc_l.data_ptr(), denc.numel(), _stream())
            d1_32, d1_16 = self._ln_bwd(d2_32, ds_q, pre1, p + "self_attn.layer_norm")
            self._dropout(d1_16, p + "self_attn.dropout", self.p_self)
            datt1 = self._lin_bwd(w["o1"], att1, d1_16, dx_f32=False)
            dqkv = torch.empty_like(qkv)
            self._att_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], datt1, dqkv[:, :hd], dqkv[:, hd:2 * hd], dqkv[:, 2 * hd:],
                          R, T, T, self_mask, T)
            de_qkv = self._lin_bwd(w["qkv"], e16_in, dqkv)
            g_a, g_b = d1_32, de_qkv
        cabi.call("cap_train_embed_bwd", tokens.data_ptr(), g_a.data_ptr(), g_b.data_ptr(), self.pad,
                  self._view(self.g32, "decoder.word_emb.components.weight").data_ptr(), Md, d, _stream())

        # ---------------- backward: encoder
        g_a, g_b = denc, None
        for l in reversed(range(L)):
            w, p = enc[l], f"encoder.layers.{l}."
            x16_in, qkv, att, pre1, y16, h16, pre2 = enc_saved[l]
            d2_32, d2_16 = self._ln_bwd(g_a, g_b, pre2, p + "pwff.layer_norm", zero_rows=enc_rows)
            self._dropout(d2_16, p + "pwff.dropout", self.p_enc)
            dh16 = self._lin_bwd(w["fc2"], h16, d2_16, dx_f32=False)
            self._dropout(dh16, p + "pwff.dropout_2", self.p_enc)
            cabi.call("cap_train_relu_bwd", dh16.data_ptr(), h16.data_ptr(), dh16.numel(), _stream())
            dy_ffn = self._lin_bwd(w["fc1"], y16, dh16)
            d1_32, d1_16 = self._ln_bwd(d2_32, dy_ffn, pre1, p + "mhatt.layer_norm")
            self._dropout(d1_16, p + "mhatt.dropout", self.p_enc)
            datt = self._lin_bwd(w["o"], att, d1_16, dx_f32=False)
            dqkv = torch.empty_like(qkv)
            self._att_bwd(qkv[:, :hd], qkv[:, hd:2 * hd], qkv[:, 2 * hd:], datt, dqkv[:, :hd], dqkv[:, hd:2 * hd], dqkv[:, 2 * hd:],
                          B, n, n, enc_mask, 0)
            dx_qkv = self._lin_bwd(w["qkv"], x16_in, dqkv)
            g_a, g_b = d1_32, dx_qkv
        _, d0_16 = self._ln_bwd(g_a, g_b, pre0, "encoder.layer_norm")
        self._dropout(d0_16, "vision_embedding.dropout", self.p_vision)
        self._lin_bwd(proj, f16, d0_16, need_dx=False)
        return loss

    def optimizer_step(self) -> None:
        """Adam + the Noam learning rate of this step (torch.optim.Adam / LambdaLR semantics)."""
        self.steps_done += 1
        lr_t = self.lr * noam_factor(self.steps_done - 1, self.d, self.warmup)
        cabi.call("cap_train_adam", self.p32.data_ptr(), self.g32.data_ptr(), self.m32.data_ptr(), self.v32.data_ptr(), self.p16.data_ptr(),
                  self.p32.numel(), lr_t, self.betas[0], self.betas[1], self.eps, self.steps_done, _stream())

    def scst_step(self, feats: Tensor, captions: Tensor, rewards: Tensor, rl_lr: float) -> Tensor:
        """One self-critical update (vi_trainer.py:121-151) given the beam search's output: ``captions`` (B, b, T) int64
        as ``model.beam_search(..., out_size=b)`` returns them, ``rewards`` (B, b) fp32 (CIDEr of each caption).

        loss = mean over the B * b beams of -(mean_T log p_t) * (r - mean_b r) (:146-148).  The log-probs the reference
        differentiates are those of the step-wise stateful decode; they equal the teacher-forced log-probs of the final
        sequences (each step attends to exactly its beam's prefix), positions after <eos> hold <pad> with log-prob 0 and
        no gradient (beam_search.py:49-55) -- so the gradient is the weighted teacher-forced backward with per-sequence
        weight advantage / (T * B * b), pinned against the reference's own backward through its beam search by
        oracle/ref_harness/gen_golden_train.py.  The optimizer of that phase is Adam(lr=RL_LEARNING_RATE), default betas,
        no scheduler (vi_trainer.py:213)."""
        B, b, T = captions.shape
        if tuple(rewards.shape) != (B, b):
            raise ValueError("rewards must be (B, beam)")
        with torch.cuda.device(self.device), torch.no_grad():
            adv = rewards.float() - rewards.float().mean(dim=1, keepdim=True)
            weights = (adv / float(T * B * b)).reshape(-1).contiguous()
            bos = torch.full((B, b, 1), self.vocab.bos_idx, dtype=torch.int64, device=captions.device)
            tokens = torch.cat([bos, captions[:, :, :-1]], dim=2).contiguous()
            loss = self.loss_and_grads(feats, tokens, captions.contiguous(), row_weights=weights)
            self.rl_steps_done = getattr(self, "rl_steps_done", 0) + 1
            if self.rl_steps_done == 1:      # a fresh optimizer: Adam's moments restart (vi_trainer.py:213)
                self.m32.zero_()
                self.v32.zero_()
            cabi.call("cap_train_adam", self.p32.data_ptr(), self.g32.data_ptr(), self.m32.data_ptr(), self.v32.data_ptr(),
                      self.p16.data_ptr(), self.p32.numel(), float(rl_lr), 0.9, 0.999, self.eps, self.rl_steps_done, _stream())
        return loss

    def step(self, feats: Tensor, tokens: Tensor, targets: Tensor) -> Tensor:
        """One iteration of vi_trainer.py:105-119; returns the loss of the batch (device tensor, no sync).

        Under ``torch.distributed`` (one process per GPU, NCCL) the arguments are this rank's shard of the batch: the
        ranks agree on the number of valid target tokens of the whole batch, weight their tokens by its inverse and sum
        the flat gradient buffers with one all-reduce -- every rank then applies the update of the GLOBAL batch, and the
        returned loss is the global mean."""
        import torch.distributed as dist
        from . import parallel
        with torch.cuda.device(self.device), torch.no_grad():
            if dist.is_initialized() and dist.get_world_size() > 1:
                inv = parallel.global_token_weight((targets != self.pad).sum())
                weights = inv.expand(tokens.shape[0] * (tokens.shape[1] if tokens.dim() == 3 else 1)).contiguous()
                loss = self.loss_and_grads(feats, tokens, targets, row_weights=weights)
                parallel.sum_gradients_(self.g32, loss)
            else:
                loss = self.loss_and_grads(feats, tokens, targets)
            self.optimizer_step()
        return loss


def self_critical_iteration(trainer: XETrainer, items, references, cider, beam_size: int, rl_lr: float,
                            use_engine: bool = False):
    """The body of ``Trainer.train_scst``'s loop (trainers/vi_trainer.py:130-151) on the native pieces: beam search on
    the engine (``out_size = beam_size``), ids -> words (``Vocab.decode_caption``), CIDEr-D reward of every beam against
    the image's references (``evaluation.Cider``), the self-critical update (``XETrainer.scst_step``).  ``use_engine=True`` samples on the whole-path engine (rebuilt
    whenever the weights changed: worth it only for large batches).

    items: InstanceList with the model's feature field (and region_boxes); references: per image, a list of reference
    captions (strings); cider: ``openviic_b200.evaluation.Cider(train references)``.
    Returns (loss, mean reward, mean baseline) -- the three numbers the reference's progress bar shows."""
    import itertools
    model, vocab = trainer.model, trainer.vocab
    trainer.sync_to_model()
    feats, _ = model.engine_inputs(items)
    bs = feats.shape[0]
    # the weights change every iteration: sample on the module-level CUDA path (its bf16 weight copies follow the
    # parameters' versions) instead of rebuilding the whole-path engine -- a state_dict round trip -- per iteration
    previous, model.disable_engine = getattr(model, "disable_engine", False), not use_engine
    try:
        outs, _ = model.beam_search(items, batch_size=bs, beam_size=beam_size, out_size=beam_size)  # (B, b, T)
    finally:
        model.disable_engine = previous
    caps_gen = vocab.decode_caption(outs.reshape(-1, outs.shape[-1]), join_words=True)
    caps_gt = list(itertools.chain(*([r] * beam_size for r in references)))
    gens = {f"{i}": [c] for i, c in enumerate(caps_gen)}
    gts = {f"{i}": r for i, r in enumerate(caps_gt)}
    reward = torch.from_numpy(cider.compute_score(gts, gens)[1].astype("float32")).to(feats.device).view(bs, beam_size)
    loss = trainer.scst_step(feats.to(trainer.device), outs.to(trainer.device), reward, rl_lr)
    return loss, reward.mean(), reward.mean(dim=1).mean()
