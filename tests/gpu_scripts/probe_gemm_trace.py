"""GPU probe: where does a GEMM CTA spend its time?  (%globaltimer phase stamps per CTA)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from openviic_b200 import cabi, ops  # noqa: E402

NAMES = ["entry", "prologue", "tma0_issued", "stage0_landed", "last_commit", "acc_visible", "stores_done", "tmem_freed"]


def main():
    dev = torch.device("cuda")
    for (m, n, k) in [(1280, 512, 512), (1280, 2048, 512), (12544, 2048, 512)]:
        x = torch.randn(m, k, device=dev).to(torch.bfloat16)
        w = torch.randn(n, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, device=dev)
        buf = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
        for _ in range(3):
            ops.linear(x, w, b)
        torch.cuda.synchronize()
        cabi.call("cap_debug_gemm_trace", buf.data_ptr())
        ops.linear(x, w, b)
        torch.cuda.synchronize()
        cabi.call("cap_debug_gemm_trace", None)
        t = buf.view(-1, 8).cpu()
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        rel = (t - t0).float() / 1e3
        print(f"== {m}x{n}x{k}: {t.shape[0]} CTAs, kernel span {float((t[:, 7].max() - t0)) / 1e3:.2f} us")
        print("   per-CTA phase offsets from the CTA's own entry (us): median / max")
        own = (t - t[:, :1]).float() / 1e3
        for i, name in enumerate(NAMES):
            print(f"   {name:14s} {own[:, i].median():7.2f} {own[:, i].max():7.2f}   (entry offset vs kernel start: median {rel[:, 0].median():.2f})" if i == 0
                  else f"   {name:14s} {own[:, i].median():7.2f} {own[:, i].max():7.2f}")


if __name__ == "__main__":
    main()
