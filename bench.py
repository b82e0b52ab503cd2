#!/usr/bin/env python
"""Headline benchmark: captions/sec, beam 5, caption length 20 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA engine)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One "step" is one pass of the hot path over one batch of synthetic visual features:
``encoder_forward`` once + 20 beam-search decode steps + (N > 1) the all-gather of caption ids.
N = 1 workload = BASELINE.json configs[1]: standard transformer, 7x7x2048 grid features, beam 5,
batch 256 on one B200, bf16.  N > 1: every rank captions its own 256 images (weak scaling), launched
by torchrun, one rank per GPU; the time is the max over ranks of the CUDA-event time on each rank.

Prints ONE JSON line (rank 0).  Keys follow the driver contract: value (device-resident inputs),
e2e (host buffers, H2D + D2H inside the timed region, through the C-ABI host entry point),
roofline (the dominant kernel family, measured live with CUDA events), cpu_baseline (the oracle port of the
reference's algorithm on this box's host cores), clocks, gpu_launches.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

RESULT_OUT = sys.stdout   # main() re-points it at the original stdout before diverting fd 1 to stderr

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
# 32 streams on the default 8 hardware work queues serialise falsely (a stream's graph launch waits behind another
# stream's queued work): measured 54 k -> 73 k captions/s.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "captions_per_sec_beam5_len20"
UNIT = "captions/s"
BEAM, MAX_LEN, VOCAB, SEED = 5, 20, 10201, 1234
WORKLOADS = {
    # name: (yaml, visual tokens, per-GPU batch, algorithmic GFLOP per caption -- BASELINE.md section 3)
    "standard_grid": ("standard_transformer.yaml", 49, 256, 4.35),
    "standard_region": ("standard_transformer_using_region.yaml", 50, 256, 4.37),
    "meshed_memory": ("meshed_memory_transformer.yaml", 50, 128, 5.97),
    "object_relation": ("object_relation_transformer.yaml", 50, 256, 4.37),
}


def measured_peaks():
    path = REPO / "MEASURED_PEAKS.json"
    if path.exists():
        p = json.loads(path.read_text())
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 50 ms from the first warm-up step to the end of the
    e2e pass (the GPU is under the same load throughout; the device-timed region alone can be < 0.2 s)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:   # N > 1: only rank 0 reports clocks; seven more pollers would only add driver-lock noise
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, sm_max, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            try:
                sm.append(float(row[0]))
                sm_max = max(sm_max, float(row[1]))
            except (ValueError, IndexError):
                continue
            for name, cell in zip(names, row[4:8]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": sm_max or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- model / inputs
def build_model(workload: str, device):
    import openviic_b200 as ov
    from openviic_b200 import synthetic
    yaml_name, n, batch, _ = WORKLOADS[workload]
    cfg = ov.get_config(yaml_name)
    cfg.MODEL.DEVICE = str(device)
    vocab = synthetic.SyntheticVocab(VOCAB, MAX_LEN)
    model = ov.build_model(cfg.MODEL, vocab).eval()
    weights = synthetic.load_synthetic_weights(model, SEED)
    return cfg, vocab, model, weights


def gemm_shapes(cfg, n: int, batch: int, levels: int):
    """(M, N, K, launches per step) of every projection GEMM in one step (encoder once + 20 decode steps)."""
    m = cfg.MODEL
    d, dff, dfeat = m.ENCODER.D_MODEL, m.ENCODER.SELF_ATTENTION.D_FF, m.VISION_EMBEDDING.D_FEATURE
    le, ld = m.ENCODER.LAYERS, m.DECODER.LAYERS
    rows, r = batch * n, batch * BEAM
    shapes = [(rows, d, dfeat, 1), (rows, 3 * d, d, le), (rows, d, d, le), (rows, dff, d, le), (rows, d, dff, le),
              (rows, 2 * d, d, ld * levels),
              (r, 3 * d, d, ld * MAX_LEN), (r, d, d, 2 * ld * MAX_LEN), (r * levels, d, d, ld * MAX_LEN),
              (r, dff, d, ld * MAX_LEN), (r, d, dff, ld * MAX_LEN), (r, VOCAB, d, MAX_LEN)]
    if levels > 1:
        shapes.append((r, d, 2 * d, ld * levels * MAX_LEN))
    return shapes


def time_gemm_family(cfg, n, batch, levels, device):
    """Average launch duration of the tcgen05 GEMM at every shape of the step, with CUDA events on the
    launching stream; returns (flops per step, seconds per step) of the whole family."""
    from openviic_b200 import ops
    total_flops, total_s = 0.0, 0.0
    for (m_, n_, k_, count) in gemm_shapes(cfg, n, batch, levels):
        x = torch.randn(m_, k_, device=device).to(torch.bfloat16)
        w = torch.randn(n_, k_, device=device).to(torch.bfloat16)
        for _ in range(3):
            ops.linear(x, w)
        torch.cuda.synchronize()
        # replay 20 back-to-back launches from a CUDA graph so the events see device time, not the
        # Python launch rate (the step itself runs from a graph too)
        reps, graph, side = 20, torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for _ in range(reps):
                    ops.linear(x, w)
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3 / (3 * reps)
        total_flops += 2.0 * m_ * n_ * k_ * count
        total_s += sec * count
    return total_flops, total_s


def time_decode_chains(eng, cfg, batch, device):
    """Average duration of the fused tcgen05 GEMM-chain launches of one decode step (1 + 2 * layers launches of
    decode_chain_kernel), with CUDA events on the launching stream around graph replays of all 20 steps'
    chains back to back; returns (algorithmic GEMM flops per step, seconds per step, launches per step) or None when
    the engine does not decode with chains."""
    import ctypes as C
    from openviic_b200 import cabi
    dec = cfg.MODEL.DECODER
    d, dff, layers = dec.D_MODEL, dec.ATTENTION.ENC_ATTENTION.D_FF, dec.LAYERS
    side = torch.cuda.Stream(device=device)
    try:
        with torch.cuda.stream(side):
            cabi.call("cap_engine_debug_chains", eng._h, 0, C.c_void_p(side.cuda_stream))
    except RuntimeError:
        return None
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for t in range(MAX_LEN):
                cabi.call("cap_engine_debug_chains", eng._h, t, C.c_void_p(side.cuda_stream))
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    sec_per_step = e0.elapsed_time(e1) / 1e3 / (3 * MAX_LEN)
    # per layer: q|k|v 3 d^2, self fc_o, cross fc_q, cross fc_o (plain: 6 d^2); the meshed decoder adds the gates' s-part
    # (levels d^2) and runs fc_o + the gate's c-part once per encoder level
    levels = dec.ATTENTION.N_ENCODER_LAYERS if dec.ARCHITECTURE == "MeshedDecoder" else 0
    proj = (5 + 3 * levels) * d * d if levels else 6 * d * d
    macs_per_row = layers * (proj + 2 * d * dff) + VOCAB * d
    return 2.0 * batch * BEAM * macs_per_row, sec_per_step, 1 + 2 * layers


def time_decode_chains_saturated(engines, streams):
    """The same chain launches with the whole GPU busy: every engine's 20 steps of chains replayed concurrently on its
    own stream (one batch alone keeps 10 of 148 SMs busy).  Returns the effective seconds per (engine, step), i.e.
    elapsed / (engines x steps), or None if anything about the extra measurement fails (it must not take the bench
    line down with it)."""
    import ctypes as C
    from openviic_b200 import cabi
    try:
        graphs = []
        for eng, st in zip(engines, streams):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(st):
                cabi.call("cap_engine_debug_chains", eng._h, 0, C.c_void_p(st.cuda_stream))
            torch.cuda.synchronize()
            with torch.cuda.stream(st):
                with torch.cuda.graph(graph, stream=st):
                    for t in range(MAX_LEN):
                        cabi.call("cap_engine_debug_chains", eng._h, t, C.c_void_p(st.cuda_stream))
            graphs.append(graph)
        reps = 3
        elapsed = None
        for timed in (False, True):   # one untimed round first
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for st in streams:
                st.wait_stream(torch.cuda.current_stream())
            for _ in range(reps if timed else 1):
                for graph, st in zip(graphs, streams):
                    with torch.cuda.stream(st):
                        graph.replay()
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
            e1.record()
            torch.cuda.synchronize()
            elapsed = e0.elapsed_time(e1) / 1e3
        return elapsed / (reps * MAX_LEN * len(engines))
    except Exception as err:   # noqa: BLE001 -- an optional extra figure
        print(f"[bench] saturated chain timing skipped: {err}", file=sys.stderr)
        return None


# ------------------------------------------------------------------------------------ CPU reference
def cpu_reference_run(workload: str, steps: int, warmup: int, sample_batch: int = 8):
    """The reference's algorithm (oracle port, fp32, as written) on this box's host cores."""
    import openviic_b200 as ov
    from openviic_b200 import synthetic
    from oracle import caption_oracle as oracle
    yaml_name, n, _, _ = WORKLOADS[workload]
    cfg = ov.get_config(yaml_name)
    cfg.MODEL.DEVICE = "cpu"
    vocab = synthetic.SyntheticVocab(VOCAB, MAX_LEN)
    model = ov.build_model(cfg.MODEL, vocab)
    weights = synthetic.load_synthetic_weights(model, SEED)
    _, feats, boxes = synthetic.synth_inputs(cfg.MODEL, sample_batch, n, SEED)
    feats = feats.to(torch.bfloat16).float()   # the values the GPU arm is handed (bf16 features), as fp32
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(warmup):
        oracle.caption_beam_search(weights, cfg.MODEL, vocab, feats, boxes, beam=BEAM)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_ids, ref_lp = oracle.caption_beam_search(weights, cfg.MODEL, vocab, feats, boxes, beam=BEAM)
    sec = time.perf_counter() - t0
    cpu_reference_run.sample = (feats, boxes, ref_ids, ref_lp)   # bench's parity line checks the GPU arm against it
    return {"value": sample_batch * steps / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} x {sample_batch} images of the same workload (beam {BEAM}, len {MAX_LEN}, V {VOCAB}, "
                      f"fp32, torch CPU ops, {sec:.1f} s)"}, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    base, sec = cpu_reference_run(args.workload, max(1, args.steps), args.warmup, args.cpu_batch)
    yaml_name, n, _, _ = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / max(1, args.steps) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{yaml_name} n={n} beam{BEAM} len{MAX_LEN} V{VOCAB}; CPU sample of "
                                   f"{args.cpu_batch} images per step", "parallelism": "host threads"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    return 0


# ---------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import torch.distributed as dist
    from openviic_b200 import cabi, parallel, synthetic

    t_start = time.perf_counter()

    poke = install_watchdog(args.hang_seconds)

    def progress(what):   # stderr breadcrumbs: a multi-rank run that stalls shows where
        poke(what)
        if rank == 0:
            print(f"[bench +{time.perf_counter() - t_start:6.1f}s] {what}", file=sys.stderr, flush=True)

    rank, world, local_rank = parallel.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the caption path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    yaml_name, n, per_gpu_batch, gflop_per_caption = WORKLOADS[args.workload]
    batch = args.batch or per_gpu_batch
    cfg, vocab, model, weights = build_model(args.workload, device)
    eng = model.engine(batch, n, BEAM)
    levels = eng.desc.n_enc_levels
    # Independent batches are pipelined over `streams` engines/streams: at batch 256 every decode kernel is a
    # single partial wave and latency-bound, so kernels of different batches co-run on the idle SMs.
    n_streams = max(1, args.streams)
    # every further engine shares the first one's device weights (cap_engine_create_shared): one 48 MB weight set that
    # stays L2-resident, its own workspaces / caches / CUDA graph
    engines = [eng] + [eng.clone() for _ in range(n_streams - 1)]
    streams = [torch.cuda.Stream(device=device) for _ in range(n_streams)]
    progress(f"{n_streams} engines ready (world {world})")
    needs_boxes = synthetic.needs_boxes(cfg.MODEL)

    # Rotating input sets: 4 x (B,n,2048) bf16 = 4 x 51 MB > 126 MB L2, so no step finds its input cached
    # (the step's own working set -- 236 MB self-KV cache + 77 MB cross K/V + 52 MB logits -- exceeds L2 too).
    n_sets = 4
    feats_host, feats_dev, boxes_host, boxes_dev = [], [], [], []
    for i in range(n_sets):
        f = synthetic.synth_features(batch, n, cfg.MODEL.VISION_EMBEDDING.D_FEATURE, SEED + 17 * i + 1000 * rank,
                                     ragged=synthetic.feature_field(cfg.MODEL) == "region_features")
        fh = f.to(torch.bfloat16).pin_memory()
        feats_host.append(fh)
        feats_dev.append(fh.to(device))
        bx = synthetic.synth_boxes(batch, n, SEED + i).pin_memory() if needs_boxes else None
        boxes_host.append(bx)
        boxes_dev.append(None if bx is None else bx.to(device))

    outs_dev = [(torch.empty((batch, 1, MAX_LEN), device=device, dtype=torch.int64),
                 torch.empty((batch, 1, MAX_LEN), device=device, dtype=torch.float32)) for _ in range(n_streams)]

    pace_stream = torch.cuda.Stream(device=device)

    def step(i):
        k = i % n_streams
        if args.pace_ms > 0:   # admission pacing: batch i may start pace_ms after batch i-1 was admitted
            with torch.cuda.stream(pace_stream):
                torch.cuda._sleep(int(args.pace_ms * 1e-3 * 1.9e9))
                ev = torch.cuda.Event()
                ev.record()
            streams[k].wait_event(ev)
        with torch.cuda.stream(streams[k]):
            return engines[k].caption_device(feats_dev[i % n_sets], boxes_dev[i % n_sets], 1, not args.no_graph,
                                             outs_dev[k])

    def final_gather():
        """N > 1: the path's only collective -- ONE all-gather of caption ids / log-probs (the last batch of every
        stream) on the timing stream, after every side stream has been joined.  It is deliberately not issued per
        step from the 32 side streams: a NCCL kernel waits for its peers on the other GPUs, and with many streams
        sharing the hardware work queues a per-stream collective can end up queued behind work that (on another
        rank) waits for it -- the 8-GPU run of the per-step variant did not finish."""
        if world == 1:
            return None
        ids = torch.stack([o[0].squeeze(1) for o in outs_dev]).reshape(-1, MAX_LEN)
        logp = torch.stack([o[1].squeeze(1) for o in outs_dev]).reshape(-1, MAX_LEN)
        return parallel.gather_captions(ids, logp, ids.shape[0] * world)

    def fork(stagger_ms: float = 0.0):
        pace_stream.wait_stream(torch.cuda.current_stream())
        """Side streams start after everything already queued on the timing stream.  With `stagger_ms`, stream k
        additionally idles k * stagger_ms first: batches that start in lock-step stay in lock-step (every engine in
        its encoder, then every engine in the same decode chain competing for the same SMs); a serving pipeline is
        staggered by its own H2D copies, as the e2e loop below is."""
        for k, st in enumerate(streams):
            st.wait_stream(torch.cuda.current_stream())
            if stagger_ms > 0 and k > 0:
                with torch.cuda.stream(st):
                    torch.cuda._sleep(int(k * stagger_ms * 1e-3 * 1.9e9))

    def join():   # the timing stream waits for every side stream
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # kernels per step, counted on one eager (non-graph) pass: graph replays launch the same kernel nodes
    eng.encode(feats_dev[0], boxes_dev[0])
    eng.beam_search(out_size=1, use_graph=False)
    torch.cuda.synchronize()
    c0 = cabi.launch_count()
    eng.encode(feats_dev[0], boxes_dev[0])
    eng.beam_search(out_size=1, use_graph=False)
    torch.cuda.synchronize()
    launches_per_step = cabi.launch_count() - c0

    clocks = ClockSampler(local_rank, enabled=rank == 0)
    clocks.__enter__()
    for k in range(1, n_streams):   # warm every engine eagerly once (sets kernel attributes before capture)
        with torch.cuda.stream(streams[k]):
            engines[k].encode(feats_dev[0], boxes_dev[0])
            engines[k].beam_search(out_size=1, use_graph=False)
    for i in range(max(3, args.warmup) * n_streams):
        step(i)
    join()
    final_gather()   # the collective is warm (communicator channels, allocator blocks) before the timed region
    progress("warm-up enqueued")
    barrier()

    def timed_pass(passes):
        """`passes` x K steps inside ONE region bracketed by barrier + synchronize, CUDA events on the timing
        stream, max over ranks.  Returns (device ms, host enqueue s, wall ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fork(args.stagger_ms)
        host_t0 = time.perf_counter()
        for i in range(args.steps * passes):
            step(i)
        enqueue_s = time.perf_counter() - host_t0
        join()
        final_gather()
        e1.record()
        barrier()
        wall = (time.perf_counter() - host_t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), enqueue_s, wall

    # K steps at ~3 ms each is a region of tens of milliseconds: launch ramp, the final gather and event jitter
    # would be a visible part of it.  One untimed calibration pass, then the K steps are repeated `passes` times
    # inside one timed region of >= --min-timed-ms; ms_per_step = region / (K x passes).
    calib_ms, _, _ = timed_pass(1)
    # 15 % on top: the calibration pass includes a little ramp-up, and the region must not end up shorter than asked
    passes = max(1, int(-(-1.15 * args.min_timed_ms // max(calib_ms, 1e-3)))) if args.min_timed_ms > 0 else 1
    if world > 1:   # every rank must run the same number of passes
        pt = torch.tensor([passes], device=device)
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        passes = int(pt.item())
    progress(f"warm-up done (calibration pass {calib_ms:.1f} ms), timing {passes} x {args.steps} steps")
    total_ms, host_enqueue_s, wall_ms = timed_pass(passes)
    steps_timed = args.steps * passes
    value = batch * world * steps_timed / (total_ms / 1e3)
    print(f"[bench] device-resident loop: {total_ms:.1f} ms on the device, host enqueue {host_enqueue_s * 1e3:.1f} ms, wall {wall_ms:.1f} ms",
          file=sys.stderr)

    # ---- e2e: host buffers through the C-ABI host entry point (H2D + compute + D2H + sync per step) ----
    outs_host = [(torch.empty((batch, 1, MAX_LEN), dtype=torch.int64).pin_memory(),
                  torch.empty((batch, 1, MAX_LEN), dtype=torch.float32).pin_memory()) for _ in range(n_streams)]

    gathered_host = (torch.empty((n_streams * batch * world, MAX_LEN), dtype=torch.int64).pin_memory(),
                     torch.empty((n_streams * batch * world, MAX_LEN), dtype=torch.float32).pin_memory())

    def e2e_step(i):
        # the C-ABI host entry point: H2D + path + D2H of this rank's captions, enqueued by one call
        k = i % n_streams
        with torch.cuda.stream(streams[k]):
            engines[k].caption_host(feats_host[i % n_sets], boxes_host[i % n_sets], 1, not args.no_graph, outs_host[k],
                                    sync=False)

    progress("device-resident loop done")
    for i in range(3 * n_streams):
        e2e_step(i)
    barrier()
    progress("e2e warm-up done, timing")
    t0 = time.perf_counter()
    for i in range(steps_timed):
        e2e_step(i)
    if world > 1:   # the final gather of the ranks' captions, from the host buffers the loop has just filled
        torch.cuda.synchronize()
        for k in range(n_streams):
            outs_dev[k][0].copy_(outs_host[k][0], non_blocking=True)
            outs_dev[k][1].copy_(outs_host[k][1], non_blocking=True)
        all_ids, all_lp = final_gather()
        gathered_host[0].copy_(all_ids, non_blocking=True)
        gathered_host[1].copy_(all_lp, non_blocking=True)
    barrier()   # every stream drained: all ids / log-probs are in host memory
    e2e_elapsed = time.perf_counter() - t0
    clocks.__exit__(None, None, None)
    e2e_s = torch.tensor([e2e_elapsed], device=device)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = batch * world * steps_timed / float(e2e_s.item())
    h2d = feats_host[0].numel() * feats_host[0].element_size() + (boxes_host[0].numel() * 4 if needs_boxes else 0)
    d2h = batch * MAX_LEN * (8 + 4)

    progress("e2e loop done")
    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel family (tcgen05 GEMM), measured live with CUDA events ----
    peaks = measured_peaks()
    chains = time_decode_chains(eng, cfg, batch, device)
    traffic = None   # dram__bytes_read+write per launch of the dominant kernel (ncu --set full capture), if committed
    if chains is not None:
        # dominant kernel: the fused GEMM chains of the decode steps (20 steps x 7 launches per batch)
        flops, chain_s, chain_launches = chains
        summary = REPO / "profiles" / "r02_chain_ncu_summary.json"
        if summary.exists() and args.workload == "standard_grid" and batch == 256:
            traffic = json.loads(summary.read_text()).get("avg_dram_bytes_per_launch")
        alone = flops / chain_s / 1e12
        sat_s = time_decode_chains_saturated(engines, streams) if n_streams > 1 else None
        # `achieved` / `frac` are the IN-REGIME figures: the chain launches of all the batches the throughput number
        # has in flight, all SMs busy.  One batch alone occupies 10 of 148 SMs; that figure is kept as `alone_*`.
        achieved = flops / sat_s / 1e12 if sat_s else alone
        roofline = {"bound": "tensor",
                    "kernel": "decode_chain_kernel (per decode step: 1 + 2*layers launches, each a job table of "
                              "projection / FFN / gate / vocabulary GEMMs with their bias, ReLU, residual + LayerNorm, "
                              "gate-mix and log-softmax-statistics epilogues, plus the token embedding)",
                    "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                    "flops_per_launch": flops / chain_launches,
                    "timing": (f"CUDA events around graph replays of the 20 steps' chain launches of {n_streams} batches "
                               f"replayed concurrently on {n_streams} streams (the regime the throughput figure runs "
                               f"in); flops of one batch / (elapsed / {n_streams})") if sat_s else
                              "CUDA events around graph replays of the 20 steps' chain launches of one batch alone",
                    "alone_achieved": alone, "alone_frac": alone / peaks["tflops_sustained"],
                    "alone_avg_launch_us": chain_s / chain_launches * 1e6,
                    "alone_timing": "the same launches of ONE batch alone on the GPU (10 row tiles = 10 of 148 SMs busy): "
                                    "per-launch latency, not throughput",
                    "share_of_step": MAX_LEN * (sat_s if sat_s else chain_s) / (total_ms / 1e3 / steps_timed)}
        # one batch alone on the GPU, the setting of the committed ncu launch list: the chains' share of THAT step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.caption_device(feats_dev[0], boxes_dev[0], 1, not args.no_graph, outs_dev[0])
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            eng.caption_device(feats_dev[0], boxes_dev[0], 1, not args.no_graph, outs_dev[0])
        e1.record()
        torch.cuda.synchronize()
        alone_step_s = e0.elapsed_time(e1) / 3e3
        roofline["alone_ms_per_step"] = alone_step_s * 1e3
        roofline["alone_share_of_step"] = MAX_LEN * chain_s / alone_step_s
    else:
        flops, gemm_s = time_gemm_family(cfg, n, batch, levels, device)
        achieved = flops / gemm_s / 1e12
        summary = REPO / "profiles" / "r01_gemm_ncu_full_summary.json"
        if summary.exists() and args.workload == "standard_grid" and batch == 256:
            traffic = json.loads(summary.read_text())["avg_dram_bytes_per_launch"]
        roofline = {"bound": "tensor", "kernel": "gemm_tn_bf16_tcgen05 (all projection/FFN/vocab GEMMs of one step)",
                    "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["tflops_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                    "share_of_step": gemm_s / (total_ms / 1e3 / steps_timed)}
    roofline["step_algorithmic_tflops"] = gflop_per_caption * 1e9 * max(value, e2e_value) / world / 1e12
    roofline["step_frac_of_tensor_peak"] = roofline["step_algorithmic_tflops"] / peaks["tflops_sustained"]

    progress("roofline timing done")
    cpu_base = None
    if not args.skip_cpu and world == 1:   # the CPU baseline is an N=1 figure (rank 0, host cores otherwise idle)
        cpu_base, _ = cpu_reference_run(args.workload, steps=args.cpu_steps, warmup=1, sample_batch=args.cpu_batch)

    parity = None
    if cpu_base is not None:   # the same sample through the GPU path (outside every timed region), against the oracle
        f32, bx, ref_ids, ref_lp = cpu_reference_run.sample
        small = model.engine(f32.shape[0], n, BEAM)   # the model's cached engine (engines[0]): the sample fits its reservation
        got_ids, got_lp = small.caption_host(f32.to(torch.bfloat16).pin_memory(), None if bx is None else bx.pin_memory(), 1,
                                             use_graph=False)
        got_ids, got_lp = got_ids.squeeze(1).cpu(), got_lp.squeeze(1).cpu()
        same = (got_ids == ref_ids).all(dim=1)
        parity = {"images": int(f32.shape[0]), "captions_identical": int(same.sum()),
                  "logprob_max_abs_on_identical": float((got_lp - ref_lp)[same].abs().max()) if same.any() else None,
                  "against": "oracle/caption_oracle.py (fp32, the reference's algorithm) on the cpu_baseline sample; the "
                             "full-size parity tests are tests/test_gpu_fullsize.py"}

    others = None
    if world == 1 and args.workload == "standard_grid" and not args.no_others and not args.batch:
        others = other_workloads(args, progress)   # the GPU is idle now; each runs as its own process

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": total_ms / steps_timed, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{yaml_name}: {n} visual tokens x 2048, beam {BEAM}, len {MAX_LEN}, V {VOCAB}, "
                               f"batch {batch}/GPU", "global_batch": batch * world, "parallelism": f"dp{world}",
                   "l2": f"{n_sets} rotating input batches ({n_sets} x {h2d / 1e6:.0f} MB > 126 MB L2); step working set > L2",
                   "timed_region": f"{args.steps} steps x {passes} passes = {steps_timed} steps in one region of {total_ms:.0f} ms "
                                   f"(>= {args.min_timed_ms:.0f} ms so that ramp-up and the final gather do not dominate)",
                   "timed_passes": passes,
                   "cuda_graph": not args.no_graph, "streams": n_streams,
                   "pipelining": f"{n_streams} independent batches in flight on {n_streams} streams/engines sharing one "
                                 "device weight set", "weights": "synthetic seed 1234 (openviic_b200/synthetic.py)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "host_features": "bf16 pinned",
                "api": "cap_engine_caption_host_async" if world == 1 else
                "cap_engine_caption_host_async per rank + one final NCCL all-gather of the caption ids"},
        "gpu_launches": int(launches_per_step * steps_timed),
        "gpu_launches_per_step": int(launches_per_step),
        "roofline": roofline, "cpu_baseline": cpu_base, "parity_sample": parity, "clocks": clocks.summary(),
        "other_workloads": others,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    return 0


def other_workloads(args, progress):
    """The same measurement for the other single-GPU configurations of BASELINE.json (the meshed-memory transformer at
    config C's per-GPU batch, the object-relation transformer of config D), each in its own process after this one has
    finished timing; their headline figures ride along in the default line (the full lines: --workload NAME)."""
    import subprocess
    out = {}
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    for name in ("meshed_memory", "object_relation"):
        progress(f"other workload: {name}")
        cmd = [sys.executable, str(Path(__file__).resolve()), "--workload", name, "--skip-cpu", "--no-others", "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--streams", str(args.streams), "--hang-seconds", str(args.hang_seconds)]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=min(240.0, max(30.0, args.hang_seconds - 5)), env=env)
            d = json.loads(res.stdout.strip().splitlines()[-1])
            out[name] = {"value": d["value"], "e2e": d["e2e"]["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                         "workload": d["config"]["workload"], "roofline_frac": d["roofline"]["frac"],
                         "step_frac_of_tensor_peak": d["roofline"]["step_frac_of_tensor_peak"],
                         "gpu_launches_per_step": d["gpu_launches_per_step"]}
        except Exception as err:   # noqa: BLE001 -- an extra figure must not take the headline line down
            out[name] = {"error": f"{type(err).__name__}: {err}"[:300]}
    return out


def install_watchdog(seconds: float):
    """A run that makes no progress for `seconds` (a device-side hang: every wait with a bound would have trapped)
    prints the Python stacks, the fault records and the flight recorder's counters, and exits: the caller -- the driver,
    torchrun, the peers of a multi-rank run -- must never be left waiting for a wedged rank."""
    import faulthandler
    import threading
    state = {"at": time.perf_counter(), "what": "start"}

    def poke(what):
        state["at"], state["what"] = time.perf_counter(), what

    def watch():
        while True:
            time.sleep(2.0)
            if time.perf_counter() - state["at"] > seconds:
                print(f"[bench] rank {os.environ.get('RANK', '0')}: no progress for {seconds:.0f} s after '{state['what']}'",
                      file=sys.stderr)
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                try:
                    from openviic_b200 import cabi
                    print(f"[bench] timed-out waits at source lines {cabi.fault_records()}; flight recorder (entered, left): "
                          f"{cabi.flight_records()}", file=sys.stderr)
                except Exception:   # noqa: BLE001
                    pass
                sys.stderr.flush()
                os._exit(4)

    threading.Thread(target=watch, daemon=True).start()
    return poke


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=320)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="standard_grid", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: the workload's)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--streams", type=int, default=32, help="independent batches kept in flight (engines/streams)")
    ap.add_argument("--stagger-ms", type=float, default=0.0,
                    help="start stream k of the device-resident loop k * this many ms late (inside the timed region)")
    ap.add_argument("--pace-ms", type=float, default=0.0,
                    help="device-resident loop: admit one batch every this many ms (0 = enqueue everything at once)")
    ap.add_argument("--min-timed-ms", type=float, default=500.0,
                    help="repeat the K steps inside the timed region until it lasts at least this long (0 = exactly K steps)")
    ap.add_argument("--hang-seconds", type=float, default=90.0,
                    help="watchdog: dump diagnostics and exit when no phase finishes within this many seconds")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true",
                    help="do not append the other single-GPU workloads' figures (other_workloads) to the default line")
    ap.add_argument("--cpu-batch", type=int, default=16)
    ap.add_argument("--cpu-steps", type=int, default=40)
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: anything a library writes to file descriptor 1 from here
    # on (NCCL prints its version banner there on rank 0) goes to stderr instead
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)
    try:
        rc = run_gpu_arm(args)
    except BaseException:   # noqa: BLE001 -- a rank that fails must not leave its peers parked in a collective
        import traceback
        traceback.print_exc()
        try:   # bounded device-side waits that timed out leave their source line in pinned host memory
            from openviic_b200 import cabi
            print(f"[bench] rank {os.environ.get('RANK', '0')}: fault records (source lines of timed-out waits): "
                  f"{cabi.fault_records()}", file=sys.stderr)
        except Exception:   # noqa: BLE001
            pass
        sys.stderr.flush()
        os._exit(1)   # no destructors, no communicator teardown: torchrun sees the exit and stops the other ranks
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
