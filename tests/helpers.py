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
    """(case, config, vocab, model, weights, field, feats, boxes) for a table entry."""
    case = CASES[name]
    cfg = apply_overrides(ov.get_config(case["config"]), case)
    cfg.MODEL.DEVICE = str(device)
    vocab = synthetic.SyntheticVocab(case["vocab"], case["max_len"])
    model = ov.build_model(cfg.MODEL, vocab).eval()
    weights = synthetic.load_synthetic_weights(model, case["seed"])
    field, feats, boxes = synthetic.synth_inputs(cfg.MODEL, case["batch"], case["n"], case["seed"])
    return case, cfg, vocab, model, weights, field, feats, boxes


def golden(name: str):
    return np.load(GOLDEN / f"{name}.npz")


def make_items(field, feats, boxes, device):
    items = ov.InstanceList()
    items.set(field, feats.to(device))
    if boxes is not None:
        items.set("region_boxes", boxes.to(device))
    return items
