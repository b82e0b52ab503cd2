"""Operator-level golden vectors from the REAL reference, for classes no runnable architecture reaches.

AdaptiveScaledDotProductAttention (models/modules/attentions.py:188-268) is only used by AdaptiveDecoder, which the
reference cannot construct -- but the attention class itself runs.  This script runs it on
synthetic inputs with synthetic weights and stores inputs' recipe + output; it refuses to write the fixture unless
oracle.adaptive_attention reproduces the class.

usage:  python oracle/ref_harness/gen_golden_ops.py
"""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path(os.environ.get("OPENVIIC_REFERENCE", "/root/reference"))
sys.path.insert(0, str(HERE / "shims"))
sys.path.insert(0, str(REFERENCE))

import models  # noqa: E402,F401
from models.modules.attentions import AdaptiveScaledDotProductAttention as RefAdaptive  # noqa: E402
from yacs.config import CfgNode  # noqa: E402

sys.path.insert(0, str(REPO))
from openviic_b200 import synthetic  # noqa: E402
from oracle import caption_oracle as oracle  # noqa: E402
from oracle.cases import ADAPTIVE_ATTENTION_CASE as CASE  # noqa: E402


def main():
    cfg = CfgNode(dict(CASE["config"]))
    torch.manual_seed(0)
    ref = RefAdaptive(cfg).eval()
    weights = synthetic.load_synthetic_weights(ref, CASE["seed"])
    q, k, sig, mask = synthetic.synth_adaptive_inputs(CASE)
    with torch.no_grad():
        out = ref(q, k, k, sig, attention_mask=mask)
        mine = oracle.adaptive_attention(weights, "", cfg, q, k, k, sig, mask)
    err = float((out - mine).abs().max())
    print(f"[adaptive_attention] oracle-vs-reference max-abs {err:.3g}")
    if err > 2e-5:
        raise SystemExit("oracle does not reproduce the reference class")
    path = REPO / "tests" / "golden" / "adaptive_attention.npz"
    np.savez_compressed(path, out=out.numpy())
    print(f"wrote {path.stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
