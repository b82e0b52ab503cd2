"""GPU probe: cap_linear (tcgen05) against torch fp32 matmul and the CUDA-core cross-check.

Run on a B200:  python tests/gpu_scripts/probe_gemm.py  (writes one JSON line per shape)
"""
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from openviic_b200.cabi import LIB_PATH  # noqa: E402

lib = C.CDLL(str(LIB_PATH))
sig = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
       C.c_void_p]
for name in ("cap_linear", "cap_linear_simt"):
    getattr(lib, name).argtypes = sig
    getattr(lib, name).restype = C.c_int
lib.cap_last_error.restype = C.c_char_p


def run(fn, x, w, bias, out_f32, act, ldy=None):
    M, K = x.shape
    N = w.shape[0]
    ldy = ldy or N
    y = torch.zeros(M, ldy, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    rc = fn(x.data_ptr(), x.stride(0), w.data_ptr(), bias.data_ptr() if bias is not None else None, y.data_ptr(), ldy,
            1 if out_f32 else 0, act, M, N, K, None)
    if rc != 0:
        raise RuntimeError(lib.cap_last_error().decode())
    torch.cuda.synchronize()
    return y[:, :N]


def main():
    torch.manual_seed(0)
    shapes = [
        (128, 32, 64, 0), (128, 128, 64, 0), (128, 64, 128, 1), (256, 512, 512, 0), (100, 200, 72, 0),
        (1280, 512, 512, 0), (1280, 1536, 512, 0), (1280, 2048, 512, 1), (1280, 512, 2048, 0),
        (12544, 512, 2048, 0), (12544, 2048, 512, 1), (1280, 10201, 512, 0), (37, 10201, 512, 0),
    ]
    ok_all = True
    for (M, N, K, act) in shapes:
        x = (torch.randn(M, K, device="cuda")).to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda") * 0.1
        ref = x.float() @ w.float().t() + bias
        if act == 1:
            ref = torch.relu(ref)
        for out_f32 in (1, 0):
            rec = {"M": M, "N": N, "K": K, "act": act, "out_f32": out_f32}
            try:
                ldy = (N + 7) // 8 * 8
                y_s = run(lib.cap_linear_simt, x, w, bias, out_f32, act, ldy)
                rec["simt_err"] = float((y_s.float() - ref).abs().max())
                y_t = run(lib.cap_linear, x, w, bias, out_f32, act, ldy)
                rec["tc_err"] = float((y_t.float() - ref).abs().max())
                rec["tc_vs_simt"] = float((y_t.float() - y_s.float()).abs().max())
                tol = 2e-3 if out_f32 else 4e-2
                rec["ok"] = bool(rec["tc_err"] < tol)
            except Exception as e:  # noqa: BLE001
                rec["error"] = str(e)
                rec["ok"] = False
            ok_all &= rec["ok"]
            print(json.dumps(rec), flush=True)
            if "error" in rec:
                return 1
    # timing of the shapes that dominate the path
    for (M, N, K) in [(12544, 512, 2048), (12544, 2048, 512), (12544, 1536, 512), (1280, 512, 512), (1280, 2048, 512),
                      (1280, 10201, 512)]:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
        ldy = (N + 7) // 8 * 8
        y = torch.empty(M, ldy, device="cuda", dtype=torch.bfloat16)
        for bn in (0,):
            for _ in range(3):
                lib.cap_linear(x.data_ptr(), K, w.data_ptr(), None, y.data_ptr(), ldy, 0, 0, M, N, K, None)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                lib.cap_linear(x.data_ptr(), K, w.data_ptr(), None, y.data_ptr(), ldy, 0, 0, M, N, K, None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            t0 = time.perf_counter()
            for _ in range(20):
                torch.nn.functional.linear(x, w)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                torch.nn.functional.linear(x, w)
            e1.record()
            torch.cuda.synchronize()
            ms_t = e0.elapsed_time(e1) / 20
            print(json.dumps({"time_shape": [M, N, K], "cap_ms": ms, "tflops": 2 * M * N * K / ms / 1e9,
                              "cublas_ms": ms_t, "cublas_tflops": 2 * M * N * K / ms_t / 1e9}), flush=True)
    print("ALL_OK" if ok_all else "SOME_FAILED")
    return 0 if ok_all else 2


if __name__ == "__main__":
    sys.exit(main())
