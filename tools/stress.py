"""Stress / determinism loop for the bench configuration (one GPU).

    python tools/stress.py [--iters 200] [--streams 32] [--workload standard_grid] [--batch 0] [--host-every 4]

Builds `streams` engines exactly as bench.py does, then replays the pipelined loop `iters` times.  Every engine
captions the SAME rotating input sets, so all outputs of an input set must be bit-identical across engines and
across iterations: any difference is a race (the kernels have no run-to-run nondeterminism by design -- no atomics,
fixed reduction orders).  A CUDA fault stops the loop and is reported with the iteration it happened in.  Prints one
JSON line; exit code 0 = clean, 1 = mismatches, 2 = CUDA fault.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from openviic_b200 import synthetic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--streams", type=int, default=32)
    ap.add_argument("--workload", default="standard_grid")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--host-every", type=int, default=4, help="every k-th iteration goes through the host-buffer entry point")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--seconds", type=float, default=0.0, help="stop after this many seconds (0 = run all iterations)")
    ap.add_argument("--hang-seconds", type=float, default=40.0, help="watchdog: give up when one iteration takes longer")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    yaml_name, n, per_gpu_batch, _ = bench.WORKLOADS[args.workload]
    batch = args.batch or per_gpu_batch
    cfg, vocab, model, weights = bench.build_model(args.workload, dev)
    eng0 = model.engine(batch, n, bench.BEAM)
    engines = [eng0] + [eng0.clone() for _ in range(args.streams - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]
    needs_boxes = synthetic.needs_boxes(cfg.MODEL)
    n_sets = 4
    feats_host, feats_dev, boxes_host, boxes_dev = [], [], [], []
    for i in range(n_sets):
        f = synthetic.synth_features(batch, n, cfg.MODEL.VISION_EMBEDDING.D_FEATURE, bench.SEED + 17 * i,
                                     ragged=synthetic.feature_field(cfg.MODEL) == "region_features")
        fh = f.to(torch.bfloat16).pin_memory()
        feats_host.append(fh)
        feats_dev.append(fh.to(dev))
        bx = synthetic.synth_boxes(batch, n, bench.SEED + i).pin_memory() if needs_boxes else None
        boxes_host.append(bx)
        boxes_dev.append(None if bx is None else bx.to(dev))
    outs_dev = [(torch.empty((batch, 1, bench.MAX_LEN), device=dev, dtype=torch.int64),
                 torch.empty((batch, 1, bench.MAX_LEN), device=dev, dtype=torch.float32)) for _ in range(args.streams)]
    outs_host = [(torch.empty((batch, 1, bench.MAX_LEN), dtype=torch.int64).pin_memory(),
                  torch.empty((batch, 1, bench.MAX_LEN), dtype=torch.float32).pin_memory()) for _ in range(args.streams)]
    for k, e in enumerate(engines):   # eager warm run (kernel attributes) before any capture
        with torch.cuda.stream(streams[k]):
            e.encode(feats_dev[0], boxes_dev[0])
            e.beam_search(out_size=1, use_graph=False)
    torch.cuda.synchronize()
    print(f"[stress] {args.streams} engines ready", file=sys.stderr, flush=True)

    golden = {}
    mism, fault, done = [], None, 0
    t0 = time.perf_counter()
    # watchdog: an iteration that does not finish within --hang-seconds is a device-side hang (a wait that is not
    # bounded, or a scheduling deadlock): report what is known and leave without waiting for the GPU
    import threading
    progress = {"iter": -1, "at": time.perf_counter()}

    def watchdog():
        while True:
            time.sleep(2.0)
            if time.perf_counter() - progress["at"] > args.hang_seconds:
                from openviic_b200 import cabi
                import faulthandler
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                print(json.dumps({"tool": "stress", "hang": True, "after_iteration": progress["iter"],
                                  "seconds_without_progress": time.perf_counter() - progress["at"],
                                  "timed_out_waits_at_source_lines": cabi.fault_records(),
                                  "flight_recorder_entered_left": cabi.flight_records(),
                                  "env": {k: v for k, v in os.environ.items() if k.startswith("OPENVIIC_")}}), flush=True)
                os._exit(3)

    threading.Thread(target=watchdog, daemon=True).start()
    try:
        for it in range(args.iters):
            host = args.host_every > 0 and it % args.host_every == args.host_every - 1
            for k, e in enumerate(engines):
                s = (it + k) % n_sets
                with torch.cuda.stream(streams[k]):
                    if host:
                        e.caption_host(feats_host[s], boxes_host[s], 1, not args.no_graph, outs_host[k], sync=False)
                    else:
                        e.caption_device(feats_dev[s], boxes_dev[s], 1, not args.no_graph, outs_dev[k])
            torch.cuda.synchronize()
            for k in range(args.streams):
                s = (it + k) % n_sets
                ids, lp = (outs_host[k] if host else outs_dev[k])
                ids, lp = ids.cpu().clone(), lp.cpu().clone()
                if s not in golden:
                    golden[s] = (ids, lp)
                    continue
                g_ids, g_lp = golden[s]
                if not torch.equal(ids, g_ids) or not torch.equal(lp.view(torch.int32), g_lp.view(torch.int32)):
                    bad = int((ids != g_ids).any(-1).sum())
                    mism.append({"iter": it, "engine": k, "set": s, "host": host, "captions_differ": bad,
                                 "logp_max_abs": float((lp - g_lp).abs().max())})
            done = it + 1
            progress["iter"], progress["at"] = it, time.perf_counter()
            if it % 20 == 0:
                print(f"[stress] iter {it}: {len(mism)} mismatches, {time.perf_counter() - t0:.1f} s", file=sys.stderr, flush=True)
            if args.seconds > 0 and time.perf_counter() - t0 > args.seconds:
                break
    except RuntimeError as err:   # a CUDA fault is sticky: report and leave
        from openviic_b200 import cabi
        fault = f"iteration {done}: {err}; timed-out waits at source lines {cabi.fault_records()}"
    out = {"tool": "stress", "workload": args.workload, "batch": batch, "streams": args.streams, "iters_done": done,
           "captions_checked": done * args.streams * batch, "mismatches": len(mism), "first_mismatches": mism[:8],
           "fault": fault, "seconds": time.perf_counter() - t0,
           "env": {k: v for k, v in os.environ.items() if k.startswith("OPENVIIC_")}}
    print(json.dumps(out), flush=True)
    if fault:
        os._exit(2)
    return 1 if mism else 0


if __name__ == "__main__":
    sys.exit(main())
