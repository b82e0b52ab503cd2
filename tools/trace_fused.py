"""Phase timing of the ONE-KERNEL-PER-STEP variant of the fused decode step (OPENVIIC_FUSED_DECODE=1,
cap_debug_fused_trace) at the bench workload; the default chain mode is profiled with ncu (tools/one_batch.py)."""
import os
import sys

os.environ["OPENVIIC_FUSED_DECODE"] = "1"
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import openviic_b200 as ov
from openviic_b200 import cabi, synthetic
from openviic_b200.engine import CaptionEngine

B, N, BEAM, T, V = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 49, 5, 20, 10201
dev = torch.device("cuda:0")
cfg = ov.get_config("standard_transformer.yaml")
cfg.MODEL.DEVICE = "cuda:0"
vocab = synthetic.SyntheticVocab(V, T)
model = ov.build_model(cfg.MODEL, vocab).eval()
synthetic.load_synthetic_weights(model, 1234)
field, feats, boxes = synthetic.synth_inputs(cfg.MODEL, B, N, 1234)
eng = CaptionEngine(cfg.MODEL, vocab, model.state_dict(), dev)
eng.reserve(B, N, BEAM)
eng.encode(feats.to(dev).bfloat16(), None)
eng.begin_decode()
tiles = (B * BEAM + 127) // 128
trace = torch.zeros(tiles * 64, dtype=torch.int64, device=dev)
names = ["wait", "embed"]
for L in range(3):
    names += [f"L{L}.qkv", f"L{L}.self", f"L{L}.ln1", f"L{L}.q", f"L{L}.cross", f"L{L}.ln2", f"L{L}.ffn1", f"L{L}.ffn2ln"]
names += ["vocab"]
for t in range(T):
    cabi.call("cap_debug_fused_trace", trace.data_ptr())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eng.decode_step(t)
    ev1.record()
    torch.cuda.synchronize()
    if t in (0, 1, 10, 19):
        tr = trace.view(tiles, 64).cpu()
        for tile in (0, tiles // 2):
            st = tr[tile, : len(names) + 1].tolist()
            d = [(st[i + 1] - st[i]) / 1e3 for i in range(len(names))]
            ep = tr[tile, 40:48].tolist()
            print("   ffn1 chunk 3 epilogue (ns): bias+sync %d, acquire %d, iters %s, release %d" % (
                ep[1] - ep[0], ep[2] - ep[1], [ep[k + 1] - ep[k] for k in range(2, 6)], ep[7] - ep[6]))
            print("   SM clock during the kernel: %.0f MHz" % ((tr[tile, 61] - tr[tile, 60]).item() / max(1, (tr[tile, 27] - tr[tile, 0]).item()) * 1e3))
            e2 = tr[tile, 48:52].tolist()
            print("   iter 1: tmem load+wait %d ns, math %d ns, staged store %d ns" % (e2[1] - e2[0], e2[2] - e2[1], ep[4] - e2[2]))
            print(f"t={t} tile={tile} step={ev0.elapsed_time(ev1)*1e3:.0f}us total={sum(d):.0f}us :: " +
                  " ".join(f"{n}={x:.1f}" for n, x in zip(names, d)))
cabi.call("cap_debug_fused_trace", None)
