"""Epilogue timing inside the default GEMM-chain decode step (cap_debug_fused_trace): %globaltimer stamps taken by
worker warp 0 around the epilogue of fc1's chunk 3 (a plain 256-column bias + ReLU + store chunk) of every row tile,
at a few decode steps of the bench workload.  One batch alone on the GPU.

    python tools/trace_chain.py [batch]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from openviic_b200 import cabi, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
cfg, vocab, model, _ = bench.build_model("standard_grid", dev)
eng = model.engine(B, 49, 5)
feats = synthetic.synth_features(B, 49, 2048, 1234, ragged=False).to(torch.bfloat16).to(dev)
eng.encode(feats, None)
eng.begin_decode()
tiles = (B * 5 + 127) // 128
trace = torch.zeros((tiles + 1) * 64, dtype=torch.int64, device=dev)
for t in range(20):
    trace.zero_()
    cabi.call("cap_debug_fused_trace", trace.data_ptr())
    eng.decode_step(t)
    torch.cuda.synchronize()
    cabi.call("cap_debug_fused_trace", None)
    if t in (0, 1, 10, 19):
        tr = trace.view(-1, 64).cpu()
        for tile in (0, tiles // 2, tiles - 1):
            ep = tr[tile, 40:48].tolist()      # 40 entry, 41 bias staged + workers' barrier, 42 accumulator acquired,
            e2 = tr[tile, 48:52].tolist()      # 43..46 end of 32-column group 0..3, 47 accumulator released
            print(f"t={t} tile={tile}: fc1 chunk 3 epilogue {ep[7] - ep[0]} ns = bias+barrier {ep[1] - ep[0]}, acquire "
                  f"{ep[2] - ep[1]}, groups {[ep[k + 1] - ep[k] for k in range(2, 6)]}, release {ep[7] - ep[6]}; "
                  f"group 1: tcgen05.ld+wait {e2[1] - e2[0]}, math {e2[2] - e2[1]}, staged store {ep[4] - e2[2]}")
