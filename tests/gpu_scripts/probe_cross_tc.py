"""GPU probe (not collected by pytest): the opt-in tensor-core decode cross-attention (OPENVIIC_CROSS_TC=1,
decode_cross_attention_tc_kernel) against the default kernel and a torch fp32 reference -- numerics at several
shapes, launch time at the bench shape, and captions / throughput of a whole beam search with the switch on.

    python tests/gpu_scripts/probe_cross_tc.py
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
    # whole path: captions with the switch on vs off (eager: the switch is read per launch; graphs bake it in)
    cfg, vocab, model, _ = bench.build_model("standard_grid", dev)
    feats = synthetic.synth_features(256, 49, 2048, 1, False).to(torch.bfloat16).to(dev)
    outs = {}
    for tensor_path in (False, True):
        os.environ["OPENVIIC_CROSS_TC"] = "1" if tensor_path else "0"
        eng = CaptionEngine(cfg.MODEL, vocab, model.state_dict(), dev)
        eng.reserve(256, 49, 5)
        ids, lp = eng.caption_device(feats, None, 1, use_graph=False)
        torch.cuda.synchronize()
        outs[tensor_path] = (ids.clone(), lp.clone())
        eng.caption_device(feats, None, 1, use_graph=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.caption_device(feats, None, 1, use_graph=True)
        e1.record()
        torch.cuda.synchronize()
        print(f"tensor_path={tensor_path}: one batch alone {e0.elapsed_time(e1) / 10:.3f} ms (graph)")
        eng.close()
    same = (outs[False][0] == outs[True][0]).all(-1).float().mean().item()
    print(f"captions identical with / without the tensor path: {same:.2%}; "
          f"max log-prob diff on identical captions "
          f"{(outs[False][1] - outs[True][1]).abs()[(outs[False][0] == outs[True][0]).all(-1)].max().item():.4f}")


if __name__ == "__main__":
    main()
