"""GPU probe: encode vs decode time, graph vs eager, for the bench workload."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench  # noqa: E402
from openviic_b200 import cabi, synthetic  # noqa: E402


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / reps * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, host


def main():
    dev = torch.device("cuda:0")
    cfg, vocab, model, _ = bench.build_model("standard_grid", dev)
    eng = model.engine(256, 49, 5)
    feats = synthetic.synth_features(256, 49, 2048, 1, False).to(torch.bfloat16).to(dev)
    eng.encode(feats)
    eng.beam_search(1, use_graph=False)
    torch.cuda.synchronize()
    print("encode        gpu ms %.3f  host ms %.3f" % timed(lambda: eng.encode(feats)))
    print("decode graph  gpu ms %.3f  host ms %.3f" % timed(lambda: eng.beam_search(1, use_graph=True)))
    print("decode eager  gpu ms %.3f  host ms %.3f" % timed(lambda: eng.beam_search(1, use_graph=False), reps=3))
    c0 = cabi.launch_count()
    eng.beam_search(1, use_graph=False)
    print("decode launches", cabi.launch_count() - c0)
    # one decode step, eager, step by step with events
    eng.begin_decode()
    per = []
    for t in range(20):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        e0.record()
        cabi.call("cap_engine_decode_logits", eng._h, t, eng._stream())
        e1.record()
        cabi.call("cap_engine_beam_advance", eng._h, t, eng._stream())
        e2.record()
        torch.cuda.synchronize()
        per.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    print("eager per-step (decoder stack ms, beam ms):", [(round(a, 3), round(b, 3)) for a, b in per[::4]])


if __name__ == "__main__":
    main()
