"""Parity at the sizes that are BENCHMARKED (BASELINE.json configs B, C per GPU, D): every image of a full batch through
the pipelined multi-engine path bench.py times, against the oracle -- the reference's fp32 arithmetic, and the same
algorithm with the CUDA path's bf16 operand rounding (oracle.operand_rounding)."""

import time

import pytest
import torch

import bench
from openviic_b200 import synthetic
from oracle import caption_oracle as oracle

pytestmark = pytest.mark.gpu

N_ENGINES = 4            # concurrent engines / streams (bench.py runs 32 of the same; 4 keep the test short)
MIN_IDENTICAL = {None: 0.80, "bf16": 0.90}   # token-identical best captions over the whole batch
TOL_CAPTION_LOGP = {None: 9e-2, "bf16": 2e-2}   # per-token log-probs of identical captions, max-abs
NEAR_TIE = 0.5           # differing captions must score within this of the oracle's caption under the oracle's own scoring


def _score(weights, cfg, vocab, feats, boxes, ids, operands):
    b = ids.shape[0]
    tokens = torch.cat([torch.full((b, 1), vocab.bos_idx, dtype=torch.long), ids[:, :-1]], 1)
    with oracle.operand_rounding(operands):
        lp = oracle.teacher_forced_log_probs(weights, cfg.MODEL, vocab, feats, tokens, boxes)
    tok = lp.gather(2, ids.unsqueeze(-1)).squeeze(-1)
    ended = (ids == vocab.eos_idx).cumsum(1) - (ids == vocab.eos_idx).long()
    return (tok * (ended == 0)).sum(1)


@pytest.mark.parametrize("workload", ["standard_grid", "meshed_memory", "object_relation"])
def test_benchmarked_configuration_matches_oracle(device, workload):
    yaml_name, n, batch, _ = bench.WORKLOADS[workload]
    cfg, vocab, model, weights = bench.build_model(workload, device)
    ragged = synthetic.feature_field(cfg.MODEL) == "region_features"
    feats = synthetic.synth_features(batch, n, cfg.MODEL.VISION_EMBEDDING.D_FEATURE, bench.SEED + 5, ragged=ragged)
    boxes = synthetic.synth_boxes(batch, n, bench.SEED + 5) if synthetic.needs_boxes(cfg.MODEL) else None
    feats16 = feats.to(torch.bfloat16)

    # ---- the pipelined path: N engines on N streams, graph replays, host buffers in and out ----
    first = model.engine(batch, n, bench.BEAM)
    engines = [first] + [first.clone() for _ in range(N_ENGINES - 1)]   # shared device weights, as in bench.py
    streams = [torch.cuda.Stream(device=device) for _ in engines]
    fh = feats16.pin_memory()
    bh = None if boxes is None else boxes.pin_memory()
    outs = None
    for rnd in range(3):   # eager, capture, replay
        outs = []
        for e, st in zip(engines, streams):
            with torch.cuda.stream(st):
                outs.append(e.caption_host(fh, bh, 1, use_graph=rnd > 0, sync=False))
        torch.cuda.synchronize()
    ids = outs[0][0].squeeze(1).clone()
    lps = outs[0][1].squeeze(1).clone()
    for o_ids, o_lp in outs[1:]:   # every engine, same input: bit-identical
        assert torch.equal(o_ids.squeeze(1), ids) and torch.equal(o_lp.squeeze(1), lps)
    for e in engines[1:]:
        e.close()

    # ---- the oracle on all images ----
    f32 = feats16.float()
    for operands in ("bf16", None):
        t0 = time.perf_counter()
        with oracle.operand_rounding(operands):
            ref_ids, ref_lp = oracle.caption_beam_search(weights, cfg.MODEL, vocab, f32, boxes, beam=bench.BEAM)
        sec = time.perf_counter() - t0
        equal = (ids == ref_ids).all(1)
        frac = equal.float().mean().item()
        err = (lps - ref_lp)[equal].abs().max().item() if equal.any() else float("nan")
        label = "bf16-operand" if operands else "fp32 reference"
        line = (f"[{workload} {batch} images] vs {label} oracle ({sec:.1f} s): captions identical {int(equal.sum())}/{batch} "
                f"= {frac:.3f}; per-token log-prob max-abs on identical captions {err:.4f}")
        if (~equal).any():
            sel = ~equal
            bx = None if boxes is None else boxes[sel]
            mine = _score(weights, cfg, vocab, f32[sel], bx, ids[sel], operands)
            theirs = _score(weights, cfg, vocab, f32[sel], bx, ref_ids[sel], operands)
            gap = theirs - mine
            line += f"; oracle-score gap of the differing captions: max {gap.max():.3f} mean {gap.mean():.3f} min {gap.min():.3f}"
            assert gap.max().item() < NEAR_TIE
        print(line)
        assert frac >= MIN_IDENTICAL[operands]
        assert not equal.any() or err < TOL_CAPTION_LOGP[operands]
