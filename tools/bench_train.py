"""Throughput of the XE training step (T1) at the caption bench's shape: config B's model, 256 images x 49 visual tokens,
captions of 20 tokens, vocabulary 10201.  CUDA events around K steps after W warm-up steps; the oracle's CPU step on a
small sample next to it.   python tools/bench_train.py [--batch 256] [--steps 10] [--cpu-batch 8]"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from openviic_b200 import cabi, synthetic  # noqa: E402
from openviic_b200.training import XETrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-batch", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg, vocab, model, weights = bench.build_model("standard_grid", dev)
    n, T, V = bench.WORKLOADS["standard_grid"][1], bench.MAX_LEN, bench.VOCAB
    trainer = XETrainer(model, lr=1.0, warmup=10000, dropout_seed=1)   # DROPOUT 0.1 as in the YAML
    sets = []
    for i in range(4):
        f = synthetic.synth_features(args.batch, n, 2048, 77 + i, ragged=False).to(torch.bfloat16).to(dev)
        tok, tgt = synthetic.synth_captions(args.batch, T, V, 77 + i)
        sets.append((f, tok.to(dev), tgt.to(dev)))
    for i in range(args.warmup):
        trainer.step(*sets[i % 4])
    torch.cuda.synchronize()
    c0 = cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = [trainer.step(*sets[i % 4]) for i in range(args.steps)]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (cabi.launch_count() - c0) // args.steps
    d, dff, L = 512, 2048, 3
    enc_tok = 2048 * d + L * (4 * d * d + 2 * d * dff)
    dec_tok = L * (4 * d * d + d * d + 2 * d * dff) + V * d       # self q|k|v|o, cross q, o (+ K|V below), FFN, vocabulary
    fwd = 2.0 * (args.batch * n * (enc_tok + L * 2 * d * d) + args.batch * T * (dec_tok + L * d * d))
    out = {"tool": "bench_train", "workload": f"standard_transformer.yaml: {args.batch} images x {n} tokens, captions of {T}, V {V}",
           "ms_per_step": ms, "images_per_s": args.batch / ms * 1e3, "launches_per_step": int(launches),
           "dropout": 0.1, "gemm_tflops_fwd_bwd": 3 * fwd / (ms * 1e-3) / 1e12, "losses": [round(x.item(), 4) for x in losses[:4]],
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if args.cpu_batch > 0:
        from oracle import caption_oracle as oracle
        cpu_feats = sets[0][0][: args.cpu_batch].float().cpu()
        tok, tgt = sets[0][1][: args.cpu_batch].cpu(), sets[0][2][: args.cpu_batch].cpu()
        torch.set_num_threads(torch.get_num_threads())
        t0 = time.perf_counter()
        oracle.xe_train_steps(weights, cfg.MODEL, vocab, [(cpu_feats, tok, tgt)] * 2, 1.0, 10000)
        sec = (time.perf_counter() - t0) / 2
        out["cpu_oracle_images_per_s"] = args.cpu_batch / sec
        out["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
