"""GPU probe (not collected by pytest): the two opt-in decode attention kernels against the default ones --
the tensor-core cross-attention (OPENVIIC_CROSS_TC=1, decode_cross_attention_tc_kernel), the split-key
self-attention (OPENVIIC_SELF_SPLIT=1, decode_self_attention_split_kernel) and the whole-image encoder self-attention
(OPENVIIC_ENC_TC=1, encoder_self_attention_tc_kernel): numerics at several shapes, launch time
at the bench shape, and captions / latency of a whole beam search with each switch on.

    python tests/gpu_scripts/probe_attention_variants.py
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from openviic_b200 import CaptionEngine, cabi, synthetic  # noqa: E402


def reference(q, kv, mask, beam, H):
    R, hd = q.shape
    B, n, _ = kv.shape
    k = kv[..., :hd].float().repeat_interleave(beam, 0).view(R, n, H, 64).transpose(1, 2)
    v = kv[..., hd:].float().repeat_interleave(beam, 0).view(R, n, H, 64).transpose(1, 2)
    s = torch.einsum("rhd,rhnd->rhn", q.float().view(R, H, 64), k) * 0.125
    s = s.masked_fill(mask.repeat_interleave(beam, 0).view(R, 1, n), float("-inf"))
    return torch.einsum("rhn,rhnd->rhd", torch.softmax(s, -1), v).reshape(R, hd)


def run(q, kv, mask, beam, H, tensor_path):
    os.environ["OPENVIIC_CROSS_TC"] = "1" if tensor_path else "0"
    B, n, _ = kv.shape
    out = torch.empty_like(q)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cabi.call("cap_decode_cross_attention", q.data_ptr(), q.shape[1], kv.data_ptr(), mask.data_ptr(), out.data_ptr(),
              q.shape[1], B, beam, n, H, 0.125, stream)
    return out


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    H, hd = 8, 512
    for B, beam, n in [(7, 5, 49), (3, 5, 50), (5, 3, 37), (4, 1, 56), (6, 5, 57), (2, 4, 99), (256, 5, 49)]:
        kv = torch.randn(B, n, 2 * hd, generator=g).to(torch.bfloat16).to(dev)
        q = torch.randn(B * beam, hd, generator=g).to(torch.bfloat16).to(dev)
        mask = torch.zeros(B, n, dtype=torch.uint8)
        mask[B // 2, n // 2:] = 1
        mask = mask.to(dev)
        ref = reference(q, kv, mask.bool(), beam, H)
        for tensor_path in (False, True):
            out = run(q, kv, mask, beam, H, tensor_path)
            torch.cuda.synchronize()
            err = (out.float() - ref).abs().max().item()
            print(f"B={B} beam={beam} n={n} tensor_path={tensor_path}: max-abs err vs fp32 reference {err:.4f}")
    # launch time at the bench shape, inputs rotated so that K|V (25.7 MB per set) does not stay in L2
    B, beam, n = 256, 5, 49
    sets = [(torch.randn(B * beam, hd, generator=g).to(torch.bfloat16).to(dev),
             torch.randn(B, n, 2 * hd, generator=g).to(torch.bfloat16).to(dev)) for _ in range(8)]
    mask = torch.zeros(B, n, dtype=torch.uint8, device=dev)
    for tensor_path in (False, True):
        for q, kv in sets:
            run(q, kv, mask, beam, H, tensor_path)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for rep in range(25):
            for q, kv in sets:
                run(q, kv, mask, beam, H, tensor_path)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (25 * len(sets))
        print(f"tensor_path={tensor_path}: {us:.2f} us per launch back to back, {B * n * 2048 / us / 1e3:.0f} GB/s of K|V")
    self_attention_part(dev, g)
    encoder_attention_part(dev, g)
    # whole path: captions with the switches on vs off (both are read per launch; CUDA graphs bake them in at capture)
    cfg, vocab, model, _ = bench.build_model("standard_grid", dev)
    feats = synthetic.synth_features(256, 49, 2048, 1, False).to(torch.bfloat16).to(dev)
    outs = {}
    for name, env in [("default", {}), ("cross_tc", {"OPENVIIC_CROSS_TC": "1"}), ("self_split", {"OPENVIIC_SELF_SPLIT": "1"}),
                      ("enc_tc", {"OPENVIIC_ENC_TC": "1"}),
                      ("all", {"OPENVIIC_CROSS_TC": "1", "OPENVIIC_SELF_SPLIT": "1", "OPENVIIC_ENC_TC": "1"})]:
        for key in ("OPENVIIC_CROSS_TC", "OPENVIIC_SELF_SPLIT", "OPENVIIC_ENC_TC"):
            os.environ[key] = env.get(key, "0")
        eng = CaptionEngine(cfg.MODEL, vocab, model.state_dict(), dev)
        eng.reserve(256, 49, 5)
        ids, lp = eng.caption_device(feats, None, 1, use_graph=False)
        torch.cuda.synchronize()
        outs[name] = (ids.clone(), lp.clone())
        eng.caption_device(feats, None, 1, use_graph=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.caption_device(feats, None, 1, use_graph=True)
        e1.record()
        torch.cuda.synchronize()
        same = (outs["default"][0] == ids).all(-1)
        print(f"{name}: one batch alone {e0.elapsed_time(e1) / 10:.3f} ms (graph); captions identical to default "
              f"{same.float().mean().item():.2%}, max log-prob diff on those "
              f"{(outs['default'][1] - lp).abs()[same].max().item():.4f}")
        eng.close()
    os.environ["OPENVIIC_CROSS_TC"] = os.environ["OPENVIIC_SELF_SPLIT"] = os.environ["OPENVIIC_ENC_TC"] = "0"


def encoder_attention_part(dev, g):
    """Whole-image encoder self-attention (OPENVIIC_ENC_TC=1) against the per-(image, head) kernel, through ops.attention
    on slices of a fused q|k|v tensor (the layout the engine's encoder uses)."""
    from openviic_b200 import ops
    H, hd = 8, 512
    for B, n in [(5, 49), (3, 50), (4, 37), (2, 57), (2, 64), (256, 49)]:
        qkv = torch.randn(B, n, 3 * hd, generator=g).to(torch.bfloat16).to(dev)
        mask = torch.zeros(B, 1, 1, n, dtype=torch.bool)
        mask[B // 2, ..., n // 2:] = True
        mask[0, ..., 0] = True
        mask = mask.to(dev)
        q, k, v = qkv[..., :hd], qkv[..., hd:2 * hd], qkv[..., 2 * hd:]
        qf, kf, vf = (x.float().view(B, n, H, 64).transpose(1, 2) for x in (q, k, v))
        sc = (qf @ kf.transpose(-1, -2)) * 0.125
        ref = (torch.softmax(sc.masked_fill(mask, float("-inf")), -1) @ vf).transpose(1, 2).reshape(B, n, hd)
        line = f"encoder attention B={B} n={n}:"
        for flag in ("0", "1"):
            os.environ["OPENVIIC_ENC_TC"] = flag
            out = ops.attention(q, k, v, H, mask=mask)
            torch.cuda.synchronize()
            line += f" enc_tc={flag} max-abs err {(out.float() - ref).abs().max().item():.4f}"
            if B == 256:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50):
                    ops.attention(q, k, v, H, mask=mask)
                e1.record()
                torch.cuda.synchronize()
                line += f" ({e0.elapsed_time(e1) * 1e3 / 50:.2f} us per call incl. host)"
        print(line)
    os.environ["OPENVIIC_ENC_TC"] = "0"


def self_attention_part(dev, g):
    H, hd, T = 8, 512, 20
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for B, beam in [(7, 5), (256, 5)]:
        R = B * beam
        qkv = torch.randn(T, R, 3 * hd, generator=g).to(torch.bfloat16).to(dev)
        anc = torch.empty(T, R, dtype=torch.int32)
        for tt in range(T):
            anc[tt] = (torch.arange(R) // beam) * beam + torch.randint(0, beam, (R,), generator=g)
        pad = torch.rand(T, R, generator=g) < 0.2
        pad[0] = False
        anc_d, pad_d = anc.to(dev), pad.to(torch.uint8).to(dev)
        for t in (0, 1, 4, 9, 19):
            outs = []
            for split in ("0", "1"):
                os.environ["OPENVIIC_SELF_SPLIT"] = split
                out = torch.empty(R, hd, dtype=torch.bfloat16, device=dev)
                cabi.call("cap_decode_self_attention", qkv.data_ptr(), anc_d.data_ptr(), pad_d.data_ptr(), out.data_ptr(),
                          hd, t, R, H, 0.125, stream)
                torch.cuda.synchronize()
                outs.append(out.float())
            line = f"self-attention B={B} t={t}: split vs default max-abs diff {(outs[0] - outs[1]).abs().max().item():.4f}"
            if B == 256:   # launch time, back to back (the 157 MB cache of one engine does not stay in L2 with 32 engines; here it may)
                for split in ("0", "1"):
                    os.environ["OPENVIIC_SELF_SPLIT"] = split
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(50):
                        cabi.call("cap_decode_self_attention", qkv.data_ptr(), anc_d.data_ptr(), pad_d.data_ptr(),
                                  out.data_ptr(), hd, t, R, H, 0.125, stream)
                    e1.record()
                    torch.cuda.synchronize()
                    line += f"; split={split} {e0.elapsed_time(e1) * 1e3 / 50:.2f} us"
            print(line)
    os.environ["OPENVIIC_SELF_SPLIT"] = "0"


if __name__ == "__main__":
    main()
