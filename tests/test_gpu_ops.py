"""Per-operator parity: every C-ABI kernel against a plain fp32 PyTorch statement of the same op."""

import math

import pytest
import torch
import torch.nn.functional as F

from openviic_b200 import ops
from oracle import caption_oracle as oracle
from helpers import TOL_ACT, TOL_F32

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("m,n,k", [(128, 32, 64), (100, 200, 72), (1280, 512, 512), (300, 1536, 512), (257, 1000, 2048),
                                   (37, 10201, 512)])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_RELU, ops.ACT_SIGMOID])
def test_linear_tcgen05_matches_fp32(device, m, n, k, act):
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n)
    x = _bf(torch.randn(m, k, generator=g)).to(device)
    w = _bf(torch.randn(n, k, generator=g) / math.sqrt(k)).to(device)
    b = (torch.randn(n, generator=g) * 0.1).to(device)
    ref = x.float() @ w.float().t() + b
    ref = {ops.ACT_NONE: ref, ops.ACT_RELU: torch.relu(ref), ops.ACT_SIGMOID: torch.sigmoid(ref)}[act]
    y32 = ops.linear(x, w, b, act=act, out_dtype=torch.float32)
    assert (y32 - ref).abs().max().item() < 2e-4          # fp32 accumulate, fp32 out
    y16 = ops.linear(x, w, b, act=act)
    assert (y16.float() - ref).abs().max().item() < TOL_ACT * max(1.0, ref.abs().max().item() / 4)
    y_simt = ops.linear(x, w, b, act=act, out_dtype=torch.float32, simt=True)
    assert (y32 - y_simt).abs().max().item() < 2e-4       # tensor-core path == CUDA-core path


def test_linear_strided_input_and_no_bias(device):
    x_full = _bf(torch.randn(64, 3 * 512)).to(device)
    x = x_full[:, 512:1024]                                # row stride 1536
    w = _bf(torch.randn(256, 512) / 22).to(device)
    y = ops.linear(x, w, None, out_dtype=torch.float32)
    assert (y - x.float() @ w.float().t()).abs().max().item() < 2e-4


@pytest.mark.parametrize("rows,d", [(1, 512), (50, 512), (1281, 512), (33, 64), (7, 2048)])
def test_add_layernorm(device, rows, d):
    g = torch.Generator().manual_seed(rows + d)
    y = torch.randn(rows, d, generator=g).to(device)
    res = _bf(torch.randn(rows, d, generator=g)).to(device)
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).to(device)
    beta = (0.1 * torch.randn(d, generator=g)).to(device)
    pos = torch.randn(5, d, generator=g).to(device)
    zero = (torch.rand(rows, generator=g) < 0.3).to(device)
    ref = F.layer_norm(y + res.float(), (d,), gamma, beta)
    out = ops.add_layernorm(y, res, gamma, beta)
    assert (out.float() - ref).abs().max().item() < TOL_ACT
    ref2 = F.layer_norm(y, (d,), gamma, beta) + pos[torch.arange(rows, device=device) % 5]
    out2 = ops.add_layernorm(y, None, gamma, beta, pos=pos)
    assert (out2.float() - ref2).abs().max().item() < TOL_ACT * 1.5
    out3 = ops.add_layernorm(y, res, gamma, beta, zero_rows=zero)
    assert torch.equal(out3[zero], torch.zeros_like(out3[zero]))
    assert torch.equal(out3[~zero], out[~zero])


@pytest.mark.parametrize("m,n,k", [(1280, 512, 512), (1280, 512, 2048), (300, 512, 512), (12544, 512, 2048), (77, 256, 64)])
def test_linear_layernorm_fused(device, m, n, k):
    """One-kernel Linear + residual + LayerNorm (cluster / DSMEM statistics) against the fp32 composition
    (attentions.py:308-309, positionwise_feed_forward.py:26) and against the two-kernel CUDA path."""
    g = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g).to(device)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(device)
    b = torch.randn(n, generator=g).to(device)
    res = torch.randn(m, n, generator=g).to(device)
    gamma = (1 + 0.1 * torch.randn(n, generator=g)).to(device)
    beta = (0.1 * torch.randn(n, generator=g)).to(device)
    zero = (torch.rand(m, generator=g) < 0.1).to(device)
    o16, o32 = ops.linear_layernorm(x, w, b, res, gamma, beta, zero_rows=zero)
    ref = torch.nn.functional.layer_norm(res + _bf(x).float() @ _bf(w).float().T + b, (n,), gamma, beta, 1e-5)
    ref = ref.masked_fill(zero[:, None], 0.0)
    assert (o32 - ref).abs().max().item() < 2e-4
    assert (o16.float() - ref).abs().max().item() < TOL_ACT
    y = ops.linear(x, w, b, out_dtype=torch.float32)
    two = ops.add_layernorm(y, res, gamma, beta, zero_rows=zero)
    assert (o32 - two).abs().max().item() < 2e-4
    # no residual, positional table
    pos = torch.randn(7, n, generator=g).to(device)
    o16, o32 = ops.linear_layernorm(x, w, b, None, gamma, beta, pos=pos)
    ref = torch.nn.functional.layer_norm(_bf(x).float() @ _bf(w).float().T + b, (n,), gamma, beta, 1e-5) + pos[torch.arange(m, device=device) % 7]
    assert (o32 - ref).abs().max().item() < 2e-4


def test_feature_mask_cast(device):
    feats = torch.randn(6, 50, 256)
    feats[1, 30:] = 0
    feats[4, 10:] = 0
    feats[5] = 0
    feats[2, 3, :2] = torch.tensor([1.5, -1.5])            # non-zero row whose partial sums cancel
    for dtype in (torch.float32, torch.bfloat16):
        f = feats.to(dtype).to(device)
        out, mask = ops.feature_mask_cast(f)
        assert torch.equal(mask.bool().cpu(), (feats.to(dtype).float().sum(-1) == 0))
        assert torch.equal(out.cpu(), feats.to(torch.bfloat16))


def _attention_ref(q, k, v, heads, mask=None, geometry=None, mem_k=None, mem_v=None):
    b, nq, hd = q.shape
    nk = k.shape[1]
    if mem_k is not None:
        k = torch.cat([k, mem_k.expand(b, -1, -1)], 1)
        v = torch.cat([v, mem_v.expand(b, -1, -1)], 1)
    qh = q.view(b, nq, heads, 64).permute(0, 2, 1, 3)
    kh = k.view(b, -1, heads, 64).permute(0, 2, 3, 1)
    vh = v.view(b, -1, heads, 64).permute(0, 2, 1, 3)
    att = qh @ kh / 8.0
    if mask is not None:
        att[:, :, :, :nk] = att[:, :, :, :nk].masked_fill(mask, -math.inf)
    if geometry is not None:
        att = att + torch.log(torch.clamp(geometry, min=1e-6))
    att = torch.softmax(att, -1)
    return (att @ vh).permute(0, 2, 1, 3).reshape(b, nq, hd)


@pytest.mark.parametrize("variant", ["sdpa", "geometry", "memory", "causal", "cross", "wide", "long_q"])
def test_attention_variants(device, variant):
    """sdpa/geometry/causal/cross: <= 64 keys and memory: 90 keys (tensor-core kernel, 8- and 16-tile forms);
    wide: 150 keys (CUDA-core kernel); long_q: 130 queries (several 64-row tiles, ragged last tile)."""
    g = torch.Generator().manual_seed(5)
    b, h, n = 5, 8, {"wide": 150, "long_q": 99}.get(variant, 50)
    nq = {"causal": 20, "cross": 7, "long_q": 130}.get(variant, n)
    nk = 20 if variant == "causal" else n
    q = _bf(torch.randn(b, nq, h * 64, generator=g))
    k = _bf(torch.randn(b, nk, h * 64, generator=g))
    v = _bf(torch.randn(b, nk, h * 64, generator=g))
    kw = {}
    mask = torch.zeros(b, 1, 1, nk, dtype=torch.bool)
    for i in range(b):
        mask[i, ..., nk - 3 * i:] = i > 0
    if variant == "causal":
        mask = mask | torch.triu(torch.ones(nq, nk), 1).bool()[None, None]
    if variant == "geometry":
        kw["geometry"] = torch.relu(torch.randn(b, h, nq, nk, generator=g))
    if variant == "memory":
        kw["mem_k"] = _bf(torch.randn(1, 40, h * 64, generator=g))
        kw["mem_v"] = _bf(torch.randn(1, 40, h * 64, generator=g))
    ref = _attention_ref(q.float(), k.float(), v.float(), h, mask, **{a: t.float() for a, t in kw.items()})
    out = ops.attention(q.to(device), k.to(device), v.to(device), h, mask=mask.to(device),
                        **{a: t.to(device) for a, t in kw.items()})
    assert (out.float().cpu() - ref).abs().max().item() < TOL_ACT


def test_attention_on_packed_qkv_views(device):
    b, n, h = 3, 49, 8
    qkv = _bf(torch.randn(b, n, 3 * h * 64)).to(device)
    q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
    out = ops.attention(q, k, v, h)
    ref = _attention_ref(q.float().cpu(), k.float().cpu(), v.float().cpu(), h)
    assert (out.float().cpu() - ref).abs().max().item() < TOL_ACT


@pytest.mark.parametrize("trig", [False, True])
def test_geometry_bias(device, trig):
    from openviic_b200.synthetic import synth_boxes
    boxes = synth_boxes(4, 50, seed=3)
    h, d_g = 8, (64 if trig else 4)
    w = torch.randn(h, d_g) * 0.5
    bias = torch.randn(h) * 0.1 + 0.1
    emb = oracle.box_relation_embedding(boxes, d_g, trig)
    ref = F.relu(torch.einsum("bijd,hd->bhij", emb, w) + bias.view(1, h, 1, 1))
    out = ops.geometry_bias(boxes.to(device), w.to(device), bias.to(device), trig).cpu()
    # sin/cos of arguments up to ~700 rad: fp32 range reduction differs slightly between libm and CUDA
    assert (out - ref).abs().max().item() < (5e-3 if trig else TOL_F32)


def test_embed_mix_gate_logsoftmax(device):
    g = torch.Generator().manual_seed(9)
    emb = _bf(torch.randn(100, 512, generator=g))
    pos = oracle.word_position_table(21, 512)
    tok = torch.randint(0, 100, (77,), generator=g)
    tok[:5] = 0
    out, flags = ops.embed_tokens(tok.to(device), emb.to(device), pos.to(device), 7, 0)
    assert (out.float().cpu() - (emb.float()[tok] + pos[7])).abs().max().item() < TOL_ACT
    assert torch.equal(flags.cpu().bool(), tok == 0)
    gates = torch.randn(3, 77, 512, generator=g)
    c = _bf(torch.randn(3, 77, 512, generator=g))
    mix = ops.meshed_mix(gates.to(device), c.to(device)).float().cpu()
    assert (mix - (torch.sigmoid(gates) * c.float()).sum(0) / math.sqrt(3)).abs().max().item() < TOL_ACT
    ig = torch.randn(77, 1024, generator=g)
    gated = ops.aoa_gate(ig.to(device)).float().cpu()
    assert (gated - ig[:, :512] * torch.sigmoid(ig[:, 512:])).abs().max().item() < TOL_ACT
    logits = torch.randn(33, 10201, generator=g) * 3
    lsm = ops.log_softmax(logits.to(device)).cpu()
    assert (lsm - F.log_softmax(logits, -1)).abs().max().item() < 1e-5


def test_decode_attention_kernels(device):
    """Beam-indirected self-attention and shared-K/V cross-attention against explicit gathers."""
    import ctypes as C
    from openviic_b200 import cabi
    g = torch.Generator().manual_seed(21)
    B, beam, H, T, t, n = 7, 5, 8, 20, 11, 49
    R, hd = B * beam, H * 64
    qkv = _bf(torch.randn(T, R, 3 * hd, generator=g)).to(device)
    anc = torch.empty(T, R, dtype=torch.int32)
    for tt in range(T):
        anc[tt] = (torch.arange(R) // beam) * beam + torch.randint(0, beam, (R,), generator=g)
    pad = (torch.rand(T, R, generator=g) < 0.2)
    pad[0] = False
    anc_d, pad_d = anc.to(device), pad.to(torch.uint8).to(device)
    out = torch.empty(R, hd, dtype=torch.bfloat16, device=device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cabi.call("cap_decode_self_attention", qkv.data_ptr(), anc_d.data_ptr(), pad_d.data_ptr(), out.data_ptr(), hd, t, R,
              H, 0.125, stream)
    qf = qkv.float().cpu()
    slots = torch.stack([anc[tt].long() if tt < t else torch.arange(R) for tt in range(t + 1)], 1)   # (R, t+1)
    steps = torch.arange(t + 1).view(1, -1).expand(R, -1)
    keys = qf[steps, slots][..., hd:2 * hd]
    vals = qf[steps, slots][..., 2 * hd:]
    q = qf[t][:, :hd].unsqueeze(1)
    mask = pad[steps, slots].view(R, 1, 1, t + 1)
    ref = _attention_ref(q, keys, vals, H, mask)
    assert (out.float().cpu() - ref.squeeze(1)).abs().max().item() < TOL_ACT

    kv = _bf(torch.randn(B, n, 2 * hd, generator=g)).to(device)
    qx = _bf(torch.randn(R, hd, generator=g)).to(device)
    kmask = torch.zeros(B, n, dtype=torch.bool)
    kmask[2, 30:] = True
    out2 = torch.empty(R, hd, dtype=torch.bfloat16, device=device)
    cabi.call("cap_decode_cross_attention", qx.data_ptr(), hd, kv.data_ptr(), kmask.to(torch.uint8).to(device).data_ptr(),
              out2.data_ptr(), hd, B, beam, n, H, 0.125, stream)
    kvf = kv.float().cpu().repeat_interleave(beam, 0)
    ref2 = _attention_ref(qx.float().cpu().unsqueeze(1), kvf[..., :hd], kvf[..., hd:], H,
                          kmask.repeat_interleave(beam, 0).view(R, 1, 1, n))
    assert (out2.float().cpu() - ref2.squeeze(1)).abs().max().item() < TOL_ACT
