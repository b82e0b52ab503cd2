"""GPU probe: per-launch time of cap_linear at the step's shapes under CUDA-graph replay (no host
overhead, warm L2), for the BLOCK_N / stage count selected by OPENVIIC_GEMM_BLOCK_N / _STAGES."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from openviic_b200 import ops  # noqa: E402

SHAPES = [(1280, 512, 512), (1280, 1536, 512), (1280, 2048, 512), (1280, 512, 2048), (1280, 10201, 512),
          (12544, 512, 2048), (12544, 1536, 512), (12544, 2048, 512), (12544, 512, 512)]


def main():
    dev = torch.device("cuda")
    out = {"bn": os.environ.get("OPENVIIC_GEMM_BLOCK_N", "auto"), "stages": os.environ.get("OPENVIIC_GEMM_STAGES", "auto")}
    for (m, n, k) in SHAPES:
        x = torch.randn(m, k, device=dev).to(torch.bfloat16)
        w = torch.randn(n, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, device=dev)
        ops.linear(x, w, b)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for _ in range(20):
                    y = ops.linear(x, w, b)
        for _ in range(2):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 100
        out[f"{m}x{n}x{k}"] = round(us, 2)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
