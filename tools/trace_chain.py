"""Where a GEMM chain's time goes (cap_debug_fused_trace): the control warps of the tracing instantiation account
their waiting time per CTA -- the tcgen05.mma issuer waiting for weight stages (TMA / producer bound), for free
accumulators (epilogue bound), for the A tile; the producer waiting for free ring stages (MMA bound) -- and worker warp
0 stamps the epilogue of fc1's fourth chunk.  One batch alone on the GPU, bench workload.

    python tools/trace_chain.py [batch] [workload]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from openviic_b200 import cabi, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
workload = sys.argv[2] if len(sys.argv) > 2 else "standard_grid"
dev = torch.device("cuda:0")
cfg, vocab, model, _ = bench.build_model(workload, dev)
n = bench.WORKLOADS[workload][1]
eng = model.engine(B, n, 5)
feats = synthetic.synth_features(B, n, 2048, 1234, ragged=False).to(torch.bfloat16).to(dev)
eng.encode(feats, None)
eng.begin_decode()
tiles = (B * 5 + 127) // 128
REGION = 64 * 64
trace = torch.zeros(3 * 6 * REGION, dtype=torch.int64, device=dev)
names = {0: "EMBED_QKV", 1: "SELF_OUT", 2: "FFN"}
for t in range(20):
    trace.zero_()
    cabi.call("cap_debug_fused_trace", trace.data_ptr())
    eng.decode_step(t)
    torch.cuda.synchronize()
    cabi.call("cap_debug_fused_trace", None)
    if t not in (1, 10, 19):
        continue
    tr = trace.view(3 * 6, 64, 64).cpu()
    for kind in (0, 1, 2):
        for layer in ((0,) if kind == 0 else (0, 2)):
            r = tr[kind * 6 + layer]
            lead = [tile for tile in range(0, tiles, 2)]          # the issuer runs in the leader CTA of each pair
            iss = r[lead][:, 0:5].float().mean(0).tolist()
            pro = r[:tiles][:, 8:11].float().mean(0).tolist()
            if iss[0] == 0:
                continue
            busy = iss[0] - sum(iss[1:5])
            print(f"t={t:2d} {names[kind]}({layer}): issuer {iss[0] / 1.965e3:7.1f} us = issue {busy / iss[0]:.0%}, wait weights "
                  f"{iss[1] / iss[0]:.0%}, wait accumulators {iss[2] / iss[0]:.0%}, wait A tile {iss[3] / iss[0]:.0%}, wait streamed A "
                  f"{iss[4] / iss[0]:.0%} | producer {pro[0] / 1.965e3:7.1f} us, waiting for free stages {pro[1] / max(pro[0], 1):.0%}, "
                  f"for hidden / A slots {pro[2] / max(pro[0], 1):.0%}")
    ep = tr[2 * 6 + 0][0, 40:48].tolist()
    if ep[0]:
        print(f"t={t:2d} FFN(0) tile 0: fc1 chunk 3 epilogue {ep[7] - ep[0]} ns = bias+barrier {ep[1] - ep[0]}, acquire "
              f"{ep[2] - ep[1]}, groups {[ep[k + 1] - ep[k] for k in range(2, 6)]}, release {ep[7] - ep[6]}")
