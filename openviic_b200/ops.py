"""Tensor-level wrappers over the C ABI: torch supplies device memory and the stream, nothing else.

Every function launches hand-written sm_100a kernels from libopenviic_cap.so on the current CUDA
stream.  Inputs must live on a CUDA device; there is no CPU path.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import cabi
from .cabi import ACT_LEAKY_RELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, CAP_BF16, CAP_F32  # noqa: F401

Tensor = torch.Tensor


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors: Optional[Tensor]) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("openviic_b200 ops run on CUDA tensors only (no CPU fallback)")


def as_bf16(x: Tensor) -> Tensor:
    return x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)


def _rows(x: Tensor) -> Tensor:
    """View (..., K) as a 2-D row-major matrix with unit inner stride."""
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and x2.stride(0) % 8 != 0):
        x2 = x2.contiguous()
    return x2


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, act: int = ACT_NONE,
           out_dtype: torch.dtype = torch.bfloat16, simt: bool = False) -> Tensor:
    """act(x @ weight.T + bias) on the tcgen05 GEMM (``simt=True``: CUDA-core cross-check)."""
    _need_cuda(x, weight, bias)
    x2 = _rows(as_bf16(x))
    w = as_bf16(weight).contiguous()
    b = None if bias is None else bias.float().contiguous()
    m, k = x2.shape
    n = w.shape[0]
    ldy = (n + 7) // 8 * 8
    y = torch.empty((m, ldy), device=x.device, dtype=out_dtype)
    cabi.call("cap_linear_simt" if simt else "cap_linear", x2.data_ptr(), x2.stride(0) if m > 1 else k, w.data_ptr(),
              _ptr(b), y.data_ptr(), ldy, CAP_F32 if out_dtype == torch.float32 else CAP_BF16, act, m, n, k, _stream())
    y = y[:, :n]
    return y.reshape(*x.shape[:-1], n)


def linear_layernorm(x: Tensor, weight: Tensor, bias: Optional[Tensor], residual: Optional[Tensor], gamma: Tensor,
                     beta: Tensor, eps: float = 1e-5, pos: Optional[Tensor] = None,
                     zero_rows: Optional[Tensor] = None):
    """LayerNorm(residual + x @ weight.T + bias) * gamma + beta in ONE kernel (cluster of N/128 CTAs per row tile
    exchanging row statistics through distributed shared memory).  Returns (bf16 copy, fp32 copy)."""
    _need_cuda(x, weight, bias, residual, gamma, beta, pos, zero_rows)
    x2 = _rows(as_bf16(x))
    w = as_bf16(weight).contiguous()
    m, k = x2.shape
    n = w.shape[0]
    b = None if bias is None else bias.float().contiguous()
    r2 = None if residual is None else residual.float().reshape(-1, n).contiguous()
    pos2 = None if pos is None else pos.float().reshape(-1, n).contiguous()
    zr = None if zero_rows is None else zero_rows.reshape(-1).to(torch.uint8).contiguous()
    out16 = torch.empty((m, n), device=x.device, dtype=torch.bfloat16)
    out32 = torch.empty((m, n), device=x.device, dtype=torch.float32)
    cabi.call("cap_linear_layernorm", x2.data_ptr(), x2.stride(0) if m > 1 else k, w.data_ptr(), _ptr(b), _ptr(r2), n,
              gamma.float().contiguous().data_ptr(), beta.float().contiguous().data_ptr(), float(eps), _ptr(pos2),
              0 if pos2 is None else pos2.shape[0], _ptr(zr), out16.data_ptr(), n, out32.data_ptr(), n, m, n, k,
              _stream())
    shape = (*x.shape[:-1], n)
    return out16.reshape(shape), out32.reshape(shape)


def add_layernorm(y: Tensor, residual: Optional[Tensor], gamma: Tensor, beta: Tensor, eps: float = 1e-5,
                  pos: Optional[Tensor] = None, zero_rows: Optional[Tensor] = None,
                  out_dtype: torch.dtype = torch.float32) -> Tensor:
    """LayerNorm(residual + y) * gamma + beta (+ pos[row % len(pos)]), rows in ``zero_rows`` zeroed.

    ``y`` / ``residual`` may be fp32 or bf16.  The result is fp32 by default: activations travel between
    modules in fp32 (as in the reference) and are rounded to bf16 only as GEMM operands."""
    _need_cuda(y, residual, gamma, beta, pos, zero_rows)
    d = y.shape[-1]

    def rows_of(t):
        t = t.reshape(-1, d)
        if t.dtype not in (torch.float32, torch.bfloat16):
            t = t.float()
        return t.contiguous()

    y2 = rows_of(y)
    r2 = None if residual is None else rows_of(residual)
    rows = y2.shape[0]
    out = torch.empty((rows, d), device=y.device, dtype=out_dtype)
    pos2 = None if pos is None else pos.float().reshape(-1, d).contiguous()
    zr = None if zero_rows is None else zero_rows.reshape(-1).to(torch.uint8).contiguous()
    code = lambda t: CAP_F32 if t.dtype == torch.float32 else CAP_BF16  # noqa: E731
    cabi.call("cap_add_layernorm", y2.data_ptr(), code(y2), d, _ptr(r2), CAP_BF16 if r2 is None else code(r2), d,
              gamma.float().contiguous().data_ptr(), beta.float().contiguous().data_ptr(), float(eps), _ptr(pos2),
              0 if pos2 is None else pos2.shape[0], _ptr(zr),
              out.data_ptr() if out_dtype == torch.bfloat16 else None, d,
              out.data_ptr() if out_dtype == torch.float32 else None, d, rows, d, _stream())
    return out.reshape(y.shape)


def feature_mask_cast(feats: Tensor):
    """(bf16 features, uint8 padding mask (B, n)) from raw fp32/bf16 features (B, n, D)."""
    _need_cuda(feats)
    if feats.dtype not in (torch.float32, torch.bfloat16):
        feats = feats.float()
    f = feats.contiguous()
    b, n, d = f.shape
    out = torch.empty((b, n, d), device=f.device, dtype=torch.bfloat16)
    mask = torch.empty((b, n), device=f.device, dtype=torch.uint8)
    cabi.call("cap_feature_mask_cast", f.data_ptr(), CAP_F32 if f.dtype == torch.float32 else CAP_BF16, out.data_ptr(),
              mask.data_ptr(), b * n, d, _stream())
    return out, mask


def geometry_bias(boxes: Tensor, w_g: Tensor, b_g: Tensor, trig: bool) -> Tensor:
    """relu(W_g . box_relation_embedding + b_g): boxes (B,n,4) -> (B,H,n,n) fp32."""
    _need_cuda(boxes, w_g, b_g)
    bx = boxes.float().contiguous()
    b, n, _ = bx.shape
    w = w_g.float().contiguous()
    h, d_g = w.shape
    g = torch.empty((b, h, n, n), device=bx.device, dtype=torch.float32)
    cabi.call("cap_geometry_bias", bx.data_ptr(), w.data_ptr(), b_g.float().contiguous().data_ptr(), g.data_ptr(), b, n,
              h, d_g, 1 if trig else 0, _stream())
    return g


def region_grid_mask(boxes: Tensor, grid_size: int) -> Tensor:
    """get_combine_masks (models/utils.py:142-154) on the device: boxes (B,n,4) in [0,1] -> bool (B,1,n,g*g), True =
    the grid cell lies outside the box's corner-to-corner cell rectangle (masked)."""
    _need_cuda(boxes)
    bx = boxes.float().contiguous()
    b, n, _ = bx.shape
    mask = torch.empty((b, n, grid_size * grid_size), device=bx.device, dtype=torch.uint8)
    cabi.call("cap_region_grid_mask", bx.data_ptr(), mask.data_ptr(), b * n, int(grid_size), _stream())
    return mask.bool().unsqueeze(1)


def _canonical_mask(mask: Optional[Tensor], b: int, nq: int, nk: int):
    """bool/uint8 mask broadcastable to (B,1,nq,nk) -> uint8 (B, nq|1, nk) + strides."""
    if mask is None:
        return None, 0, 0
    m = mask
    if m.dim() == 4:
        if m.shape[1] != 1:
            raise RuntimeError("attention masks must be head-independent (shape (B,1,nq|1,nk))")
        m = m[:, 0]
    if m.dim() != 3:
        raise RuntimeError(f"unsupported attention mask shape {tuple(mask.shape)}")
    q_rows = m.shape[1]
    if q_rows not in (1, nq) or m.shape[2] != nk or m.shape[0] not in (1, b):
        raise RuntimeError(f"attention mask {tuple(mask.shape)} does not broadcast to ({b},1,{nq},{nk})")
    m = m.expand(b, q_rows, nk).to(torch.uint8).contiguous()
    return m, q_rows * nk, (nk if q_rows == nq and nq > 1 else 0)


def attention(q: Tensor, k: Tensor, v: Tensor, heads: int, mask: Optional[Tensor] = None,
              geometry: Optional[Tensor] = None, mem_k: Optional[Tensor] = None, mem_v: Optional[Tensor] = None,
              scale: Optional[float] = None, sentinel: Optional[Tensor] = None) -> Tensor:
    """Fused multi-head softmax(q k^T * scale [+mask] [+log g] [| memory] [| sentinel]) v with d_k = d_v = 64.

    q (B,nq,H*64), k/v (B,nk,H*64) bf16 (row-strided views are fine); returns (B,nq,H*64) bf16.
    ``sentinel`` (B,nq,H*64): query i attends to one extra, never-masked key = value sentinel[:, i]
    (AdaptiveScaledDotProductAttention's language signal, attentions.py:250-263).
    """
    _need_cuda(q, k, v, mask, geometry, mem_k, mem_v, sentinel)
    q, k, v = as_bf16(q), as_bf16(k), as_bf16(v)
    b, nq, hd = q.shape
    nk = k.shape[1]
    if hd != heads * 64:
        raise RuntimeError("attention kernels are specialised for head dim 64")

    def strided(t):
        if t.stride(-1) != 1 or t.stride(1) % 2 != 0:
            t = t.contiguous()
        return t

    q, k, v = strided(q), strided(k), strided(v)
    out = torch.empty((b, nq, hd), device=q.device, dtype=torch.bfloat16)
    m, m_bs, m_qs = _canonical_mask(mask, b, nq, nk)
    geo = None if geometry is None else geometry.float().contiguous()
    mk = None if mem_k is None else as_bf16(mem_k).reshape(-1, hd).contiguous()
    mv = None if mem_v is None else as_bf16(mem_v).reshape(-1, hd).contiguous()
    sn = None
    if sentinel is not None:
        if tuple(sentinel.shape) != (b, nq, hd):
            raise RuntimeError(f"sentinel must be {(b, nq, hd)}, got {tuple(sentinel.shape)}")
        sn = strided(as_bf16(sentinel))
    args = cabi.AttentionArgs(
        q=q.data_ptr(), k=k.data_ptr(), v=v.data_ptr(), out=out.data_ptr(),
        q_bs=q.stride(0), k_bs=k.stride(0), v_bs=v.stride(0), o_bs=out.stride(0),
        ldq=q.stride(1), ldk=k.stride(1), ldv=v.stride(1), ldo=out.stride(1),
        mask=_ptr(m), mask_bs=m_bs, mask_qs=m_qs, geometry=_ptr(geo), mem_k=_ptr(mk), mem_v=_ptr(mv),
        n_mem=0 if mk is None else mk.shape[0], B=b, H=heads, nq=nq, nk=nk,
        scale=float(scale if scale is not None else 1.0 / math.sqrt(64)),
        sentinel=_ptr(sn), s_bs=0 if sn is None else sn.stride(0), lds=0 if sn is None else sn.stride(1))
    cabi.call("cap_attention", C.byref(args), _stream())
    return out


def embed_tokens(tokens: Tensor, word_emb: Tensor, pos_table: Tensor, position: int, pad_idx: int):
    """(Emb[token] + pos_table[position]) bf16 (R,d) and the uint8 pad flags (R,)."""
    _need_cuda(tokens, word_emb, pos_table)
    tok = tokens.reshape(-1).to(torch.int32).contiguous()
    emb = as_bf16(word_emb).contiguous()
    r, d = tok.shape[0], emb.shape[1]
    out = torch.empty((r, d), device=tok.device, dtype=torch.bfloat16)
    flags = torch.empty((r,), device=tok.device, dtype=torch.uint8)
    cabi.call("cap_embed_tokens", tok.data_ptr(), emb.data_ptr(), pos_table.float().contiguous().data_ptr(),
              int(position), int(pad_idx), out.data_ptr(), None, flags.data_ptr(), r, d, _stream())
    return out, flags


def meshed_mix(gates: Tensor, c: Tensor) -> Tensor:
    """sum_i sigmoid(gates[i]) * c[i] / sqrt(levels); gates fp32, c fp32/bf16, both (levels, R, d) -> fp32."""
    _need_cuda(gates, c)
    g = gates.float().contiguous()
    cc = (c if c.dtype in (torch.float32, torch.bfloat16) else c.float()).contiguous()
    levels, r, d = cc.shape[0], cc[0].numel() // cc.shape[-1], cc.shape[-1]
    out = torch.empty(cc.shape[1:], device=cc.device, dtype=torch.float32)
    cabi.call("cap_meshed_mix", g.data_ptr(), cc.data_ptr(), CAP_F32 if cc.dtype == torch.float32 else CAP_BF16, None,
              out.data_ptr(), levels, r, d, _stream())
    return out


def aoa_gate(info_gate: Tensor) -> Tensor:
    """info * sigmoid(gate) for fp32 (..., 2d) = (info | gate) -> fp32 (..., d)."""
    _need_cuda(info_gate)
    ig = info_gate.float().contiguous()
    d = ig.shape[-1] // 2
    rows = ig.numel() // (2 * d)
    out = torch.empty((*ig.shape[:-1], d), device=ig.device, dtype=torch.float32)
    cabi.call("cap_aoa_gate", ig.data_ptr(), None, out.data_ptr(), rows, d, _stream())
    return out


def log_softmax(logits: Tensor) -> Tensor:
    """Row-wise log-softmax over the last dim of an fp32 tensor (row-strided views are fine)."""
    _need_cuda(logits)
    v = logits.shape[-1]
    x = logits.float()
    x2 = x.reshape(-1, v)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    out = torch.empty((x2.shape[0], v), device=x.device, dtype=torch.float32)
    cabi.call("cap_log_softmax", x2.data_ptr(), x2.stride(0) if x2.shape[0] > 1 else v, out.data_ptr(), v, x2.shape[0], v,
              _stream())
    return out.reshape(logits.shape)


def _param_cache(kind: str, params, build):
    """Derived tensor (bf16 cast / stacked bias) cached ON the first parameter object, so it dies with
    the parameter; invalidated when any source is modified in place, replaced or moved."""
    key = (kind,) + tuple((id(p), p.data_ptr(), p._version, str(p.device)) for p in params)
    store = params[0].__dict__.setdefault("_cap_cache", {})
    hit = store.get(kind + str(len(params)))
    if hit is not None and hit[0] == key:
        return hit[1]
    with torch.no_grad():
        value = build()
    store[kind + str(len(params))] = (key, value)
    return value


def cached_bf16(*params: Tensor) -> Tensor:
    """bf16 copy of one parameter, or the row-wise stack of several (e.g. fc_q|fc_k|fc_v)."""
    def build():
        value = torch.cat([p.detach() for p in params], dim=0) if len(params) > 1 else params[0].detach()
        return value.to(torch.bfloat16).contiguous()
    return _param_cache("bf16", params, build)


def cached_f32_cat(*params: Tensor) -> Tensor:
    """fp32 concatenation of several 1-D parameters (stacked biases)."""
    return _param_cache("f32cat", params,
                        lambda: torch.cat([p.detach().float().reshape(-1) for p in params], dim=0).contiguous())
