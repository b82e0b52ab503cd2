"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

import openviic_b200 as ov
from openviic_b200 import synthetic
from oracle import caption_oracle as oracle
from oracle.cases import CASES, apply_overrides

GOLDEN = Path(__file__).resolve().parent / "golden"

# Parity tolerances: bf16 activations with fp32 accumulation.
TOL_ACT = 2e-2     # max-abs on attention outputs / encoder features / logits / log-probs, bf16
TOL_F32 = 1e-4     # kernels whose arithmetic is fp32 end to end


def load_case(name: str, device="cpu"):
    """(case, config, vocab, model, weights, field, feats, boxes) for a table entry.  Dual-path cases (a "grid" key):
    field is None, feats = (region features, grid features), boxes = (region boxes, grid boxes) -- the oracle's
    calling convention for them; make_items() turns them into the four InstanceList fields."""
    case = CASES[name]
    cfg = apply_overrides(ov.get_config(case["config"]), case)
    cfg.MODEL.DEVICE = str(device)
    vocab = synthetic.SyntheticVocab(case["vocab"], case["max_len"])
    model = ov.build_model(cfg.MODEL, vocab).eval()
    weights = synthetic.load_synthetic_weights(model, case["seed"])
    if "grid" in case:
        x = synthetic.synth_dual_inputs(cfg.MODEL, case["batch"], case["n"], case["grid"], case["seed"])
        return (case, cfg, vocab, model, weights, None, (x["region_features"], x["grid_features"]),
                (x["region_boxes"], x["grid_boxes"]))
    field, feats, boxes = synthetic.synth_inputs(cfg.MODEL, case["batch"], case["n"], case["seed"])
    return case, cfg, vocab, model, weights, field, feats, boxes


def golden(name: str):
    return np.load(GOLDEN / f"{name}.npz")


def make_items(field, feats, boxes, device):
    items = ov.InstanceList()
    if field is None:   # dual-path inputs
        items.set("region_features", feats[0].to(device))
        items.set("grid_features", feats[1].to(device))
        items.set("region_boxes", boxes[0].to(device))
        items.set("grid_boxes", boxes[1].to(device))
        return items
    items.set(field, feats.to(device))
    if boxes is not None:
        items.set("region_boxes", boxes.to(device))
    return items


# ---------------------------------------------------------------------------------------------------------------
# Step-wise comparison of an engine with the oracle: full-vocabulary log-probs while the beams agree, and a NEAR-TIE
# proof at the first step where they stop agreeing.  Beam search is a greedy filter: once two runs keep different
# prefixes their final captions (and final scores) may differ by any amount, so "the captions differ" is explained
# only at the point of divergence -- there, under the ORACLE's own candidate scores, the engine's k-th pick must be
# within a tolerance of the oracle's k-th pick (a swap of near-equal candidates, or a near-equal candidate at the
# cut-off).  Images that never diverge produce the oracle's captions token for token.
# ---------------------------------------------------------------------------------------------------------------
def stepwise_against_oracle(eng, weights, model_cfg, vocab, feats, boxes, beam, device, operands=None):
    """Returns a dict: ids / lps (engine, best caption), ref_ids / ref_lp, worst / worst_mean (max / mean abs log-prob
    difference over the full vocabulary, worst step, images still in agreement), compared (steps), agree (B,) bool,
    margins (list of (image, step, margin)) -- the near-tie margin at each image's first divergence."""
    b = feats.shape[0]
    T = vocab.max_caption_length
    trace, ltrace = [], []
    with oracle.operand_rounding(operands):
        ref_ids, ref_lp = oracle.caption_beam_search(weights, model_cfg, vocab, feats, boxes, beam=beam, out_size=1,
                                                     trace=trace, logits_trace=ltrace)
    eng.encode(feats.to(device), None if boxes is None else boxes.to(device))
    eng.begin_decode()
    agree = torch.ones(b, dtype=torch.bool)
    worst, worst_mean, compared = 0.0, 0.0, 0
    margins = []
    omask = torch.ones(b, beam)                       # the oracle's seq_mask (beam_search.py:49-51)
    prev_words = None
    for t in range(T):
        cur = 1 if t == 0 else beam
        logits = eng.decode_logits(t)
        lp = torch.log_softmax(logits.float(), -1).cpu().view(b, beam, -1)
        ref = ltrace[t].view(b, cur, -1)
        if agree.any():
            mine = lp[:, :1] if t == 0 else lp
            live = agree.clone()
            diff = (mine - ref).abs()
            if t > 0:   # rows fed <pad> (finished beams) produce a zeroed hidden state on both sides: skip them
                fed_pad = (prev_words == vocab.padding_idx)
                diff = diff.masked_fill(fed_pad.unsqueeze(-1), 0)
            d = diff[live]
            worst = max(worst, d.max().item())
            worst_mean = max(worst_mean, d.mean().item())
            compared += 1
        eng.beam_advance(t)
        parents = eng.beam_parents().cpu().view(b, beam).long()
        tokens = eng.beam_tokens().cpu().view(b, beam).long()
        # the oracle's candidate scores at this step (beam_search.py:44-55)
        if t > 0:
            omask = omask * (prev_words != vocab.eos_idx).float()
        seq_lp = torch.zeros(b, 1) if t == 0 else trace[t - 1]["seq_logprob"]
        same = (parents == trace[t]["beam"]).all(1) & (tokens == trace[t]["word"]).all(1)
        newly = agree & ~same
        for i in newly.nonzero().flatten().tolist():
            picks = []
            for k in range(beam):
                p, w = int(parents[i, k]), int(tokens[i, k])
                if t == 0:
                    p = 0
                if omask[i, p] > 0:
                    picks.append(float(seq_lp[i, p if t > 0 else 0] + ref[i, p, w]))
                else:
                    picks.append(float(seq_lp[i, p]) if w == 0 else oracle.NEG_SENTINEL)
            theirs = trace[t]["seq_logprob"][i].tolist()
            margins.append((i, t, max(o - m for o, m in zip(theirs, picks))))
        agree &= same
        omask = torch.gather(omask, 1, trace[t]["beam"])
        prev_words = trace[t]["word"]
    ids, lps = eng.finalize(1)
    torch.cuda.synchronize()
    ids, lps = ids.squeeze(1).cpu(), lps.squeeze(1).cpu()
    return dict(ids=ids, lps=lps, ref_ids=ref_ids, ref_lp=ref_lp, equal=(ids == ref_ids).all(1), agree=agree, worst=worst,
                worst_mean=worst_mean, compared=compared, margins=margins)


def bf16_operand_yardstick(weights, model_cfg, vocab, feats, boxes, beam, ref_ids):
    """How far does the reference ALGORITHM move when only its GEMM operands are rounded to bf16 (the precision the
    CUDA path computes in)?  Returns (fraction of best captions identical to the fp32 oracle's, max-abs difference of
    the teacher-forced full-vocabulary log-probs along the fp32 oracle's captions)."""
    with oracle.operand_rounding("bf16"):
        ids16, _ = oracle.caption_beam_search(weights, model_cfg, vocab, feats, boxes, beam=beam, out_size=1)
    b = ref_ids.shape[0]
    tokens = torch.cat([torch.full((b, 1), vocab.bos_idx, dtype=torch.long), ref_ids[:, :-1]], 1)
    lp32 = oracle.teacher_forced_log_probs(weights, model_cfg, vocab, feats, tokens, boxes)
    with oracle.operand_rounding("bf16"):
        lp16 = oracle.teacher_forced_log_probs(weights, model_cfg, vocab, feats, tokens, boxes)
    live = (tokens != vocab.padding_idx)
    return (ids16 == ref_ids).all(1).float().mean().item(), (lp16 - lp32).abs()[live].max().item()
