"""GPU probe: phase stamps of the vocabulary GEMM with the chunk-statistics epilogue vs the plain store epilogue."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from openviic_b200 import cabi, ops  # noqa: E402

NAMES = ["entry", "prologue", "tma0_issued", "stage0_landed", "last_commit", "acc_visible", "stores_done", "tmem_freed"]


def report(tag, buf):
    t = buf.view(-1, 8).cpu()
    t = t[t[:, 0] > 0]
    own = (t - t[:, :1]).float() / 1e3
    span = float(t[:, 7].max() - t[:, 0].min()) / 1e3
    print(f"== {tag}: {t.shape[0]} CTAs, kernel span {span:.2f} us; per-CTA median offsets (us): " +
          ", ".join(f"{n}={own[:, i].median():.2f}" for i, n in enumerate(NAMES)))


def main():
    dev = torch.device("cuda")
    R, V, d = 1280, 10201, 512
    x = torch.randn(R, d, device=dev).to(torch.bfloat16)
    w = (torch.randn(V, d, device=dev) * 0.13).to(torch.bfloat16)
    ld = (V + 7) // 8 * 8
    chunks = ((V + 255) // 256) * 8
    logits = torch.empty(R, ld, device=dev)
    pm = torch.empty(R, chunks, 2, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nc = C.c_int()

    def stats():
        cabi.call("cap_vocab_logits_stats", x.data_ptr(), d, w.data_ptr(), None, logits.data_ptr(), ld, R, V, d,
                  pm.data_ptr(), C.byref(nc), stream)

    for fn, tag in ((stats, "chunk-stats epilogue"), (lambda: ops.linear(x, w, None, out_dtype=torch.float32), "fp32 store epilogue")):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        buf = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
        cabi.call("cap_debug_gemm_trace", buf.data_ptr())
        fn()
        torch.cuda.synchronize()
        cabi.call("cap_debug_gemm_trace", None)
        report(tag, buf)


if __name__ == "__main__":
    main()
