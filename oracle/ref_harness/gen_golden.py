"""Generate tests/golden/*.npz by running the REAL reference (from /root/reference) on CPU.

Runs only in the build container (the reference cannot travel to the GPU box).  For every case it

  1. imports the reference with two import shims (yacs, termcolor -- absent in this image),
  2. builds the reference model from this repo's YAML (same schema as the reference's),
  3. overwrites its parameters with ``openviic_b200.synthetic`` weights (regenerable anywhere
     from name/shape/seed) and feeds synthetic features,
  4. records the reference's outputs (beam ids/log-probs, encoder output, step-0 log-probs,
     teacher-forced log-probs -- sub-sampled to stay small),
  5. runs ``oracle/caption_oracle.py`` on the same inputs and REFUSES to write the fixture unless
     the oracle reproduces the reference (ids identical, floats within 2e-5).

Documented oracle patch: ObjectRelationTransformer.encoder_forward passes an
``Instance`` to an encoder that wants keyword tensors (object_relation_transformer.py:38-42 vs
encoders.py:93) and raises TypeError as shipped; the harness calls the encoder with
``features= / boxes= / padding_mask=`` instead.  Nothing else in the reference is modified.

usage:  python oracle/ref_harness/gen_golden.py [case ...]
"""

from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path(os.environ.get("OPENVIIC_REFERENCE", "/root/reference"))
sys.path.insert(0, str(HERE / "shims"))
sys.path.insert(0, str(REFERENCE))

import models  # noqa: E402,F401  (reference package: fills the reference's registries)
from builders.model_builder import build_model as ref_build_model  # noqa: E402
from configs.utils import get_config as ref_get_config  # noqa: E402
from models.object_relation_transformer import ObjectRelationTransformer  # noqa: E402
from utils.instance import InstanceList as RefInstanceList  # noqa: E402

sys.path.insert(0, str(REPO))
from openviic_b200 import synthetic  # noqa: E402
from openviic_b200.configs import get_config  # noqa: E402
from oracle import caption_oracle as oracle  # noqa: E402
from oracle.cases import CASES, apply_overrides  # noqa: E402


def _ort_encoder_forward(self, input_features):
    feats, mask = self.vision_embedding(input_features.region_features)
    enc = self.encoder(features=feats, boxes=input_features.region_boxes, padding_mask=mask)
    return enc, mask


ObjectRelationTransformer.encoder_forward = _ort_encoder_forward  # the documented call-site patch


def run_case(name: str, case: dict) -> dict:
    cfg_path = REPO / "openviic_b200" / "configs" / case["config"]
    ref_cfg = apply_overrides(ref_get_config(str(cfg_path)), case)
    ref_cfg.MODEL.DEVICE = "cpu"
    vocab = synthetic.SyntheticVocab(case["vocab"], case["max_len"])
    torch.manual_seed(0)
    model = ref_build_model(ref_cfg.MODEL, vocab).eval()
    weights = synthetic.load_synthetic_weights(model, case["seed"])
    field, feats, boxes = synthetic.synth_inputs(ref_cfg.MODEL, case["batch"], case["n"], case["seed"])
    items = RefInstanceList()
    items.set(field, feats)
    if boxes is not None:
        items.set("region_boxes", boxes)

    b, beam = case["batch"], case["beam"]
    with torch.no_grad():
        t0 = time.perf_counter()
        ids, logp = model.beam_search(items, batch_size=b, beam_size=beam, out_size=1)
        ref_seconds = time.perf_counter() - t0
        ids_all, logp_all = model.beam_search(items, batch_size=b, beam_size=beam, out_size=beam)
        enc, enc_mask = model.encoder_forward(items)
        tf_tokens = torch.cat([torch.full((b, 1), vocab.bos_idx, dtype=torch.long), ids[:, :-1]], dim=1)
        items.set("caption_tokens", tf_tokens)
        tf_logp = model(items)

    # ---- the oracle must reproduce the reference before the fixture is trusted ----
    our_cfg = apply_overrides(get_config(cfg_path), case)
    logits_trace = []
    o_ids, o_logp = oracle.caption_beam_search(weights, our_cfg.MODEL, vocab, feats, boxes, beam=beam, out_size=1,
                                               logits_trace=logits_trace)
    o_ids_all, o_logp_all = oracle.caption_beam_search(weights, our_cfg.MODEL, vocab, feats, boxes, beam=beam,
                                                       out_size=beam)
    with torch.no_grad():
        o_enc, o_mask = oracle.encode(weights, our_cfg.MODEL, feats, boxes)
    o_tf = oracle.teacher_forced_log_probs(weights, our_cfg.MODEL, vocab, feats, tf_tokens, boxes)
    checks = {
        "ids": bool(torch.equal(ids, o_ids)), "ids_all": bool(torch.equal(ids_all, o_ids_all)),
        "mask": bool(torch.equal(enc_mask, o_mask)),
        "logp": float((logp - o_logp).abs().max()), "logp_all": float((logp_all - o_logp_all).abs().max()),
        "enc": float((enc - o_enc).abs().max()), "tf": float((tf_logp - o_tf).abs().max()),
    }
    ok = checks["ids"] and checks["ids_all"] and checks["mask"] and max(
        checks["logp"], checks["logp_all"], checks["enc"], checks["tf"]) < 2e-5
    print(f"[{name}] reference {ref_seconds:.2f}s  oracle-vs-reference {checks}  {'OK' if ok else 'MISMATCH'}")
    if not ok:
        raise SystemExit(f"oracle does not reproduce the reference on case {name}")

    enc_flat = enc.reshape(-1, enc.shape[-1])
    fixture = {
        "ids": ids.numpy(), "logp": logp.numpy(), "ids_all": ids_all.numpy(), "logp_all": logp_all.numpy(),
        "enc_mask": enc_mask.reshape(b, -1).numpy(),
        "enc_rows": enc_flat[:: max(1, enc_flat.shape[0] // 64)][:64].numpy(),
        "enc_row_stride": np.int64(max(1, enc_flat.shape[0] // 64)),
        "step0_logp": logits_trace[0][:, :: max(1, case["vocab"] // 256)].numpy(),
        "tf_logp": tf_logp[:, :, :: max(1, case["vocab"] // 128)].numpy(),
        "n_eos": np.int64(int((ids == vocab.eos_idx).sum())),
        "ref_seconds": np.float64(ref_seconds),
    }
    return fixture


def main(argv):
    names = argv or list(CASES)
    out_dir = REPO / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    for name in names:
        fixture = run_case(name, CASES[name])
        np.savez_compressed(out_dir / f"{name}.npz", **fixture)
        size = (out_dir / f"{name}.npz").stat().st_size
        print(f"[{name}] wrote {size / 1024:.0f} KiB, captions ending in <eos>: {int(fixture['n_eos'])}/{len(fixture['ids'])}")


if __name__ == "__main__":
    main(sys.argv[1:])
