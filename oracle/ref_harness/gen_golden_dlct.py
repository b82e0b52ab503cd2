"""Golden fixture of the dual-path (DLCT) model: the REAL reference modules, composed and repaired.

The reference ships GeometricDualFeatureEmbedding (vision_embeddings.py:45-71) and DualCollaborativeLevelEncoder
(encoders.py:114-211) but no architecture that uses them, and neither runs as written.  This
harness builds them with the reference's own builders from this repo's dlct_transformer.yaml, wires them to the
reference's Decoder through the reference's BaseTransformer, and applies the three repairs P1-P3 documented at the top
of the dual-path section of oracle/caption_oracle.py -- by wrapping reference functions, not by re-typing them:

  P1  models.modules.vision_embeddings.get_combine_masks  -> result.squeeze(1)
  P2  GeometricDualFeatureEmbedding.forward               -> the two torch.cat calls see key-padding masks expanded
                                                              over the query dim (a 12-line body of ours around the
                                                              module's own region_proj / grid_proj and the reference's
                                                              generate_padding_mask / get_combine_masks)
  P3  EncoderLayer.forward                                 -> when it is handed a (B,1,nq,nk) mask as `padding_mask`,
                                                              the query stream's own padding mask is used for the row
                                                              zeroing (picked by the number of queries)

It then refuses to write the fixture unless oracle/caption_oracle.py reproduces the composition exactly.

usage:  python oracle/ref_harness/gen_golden_dlct.py
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

import models  # noqa: E402,F401  (reference package)
import models.modules.vision_embeddings as ref_ve  # noqa: E402
from builders.decoder_builder import build_decoder as ref_build_decoder  # noqa: E402
from builders.encoder_builder import build_encoder as ref_build_encoder  # noqa: E402
from builders.vision_embedding_builder import build_vision_embedding as ref_build_vision_embedding  # noqa: E402
from configs.utils import get_config as ref_get_config  # noqa: E402
from models.base_transformer import BaseTransformer as RefBaseTransformer  # noqa: E402
from models.modules.encoders import EncoderLayer as RefEncoderLayer  # noqa: E402
from utils.instance import InstanceList as RefInstanceList  # noqa: E402

sys.path.insert(0, str(REPO))
from openviic_b200 import synthetic  # noqa: E402
from openviic_b200.configs import get_config  # noqa: E402
from oracle import caption_oracle as oracle  # noqa: E402
from oracle.cases import CASES, apply_overrides  # noqa: E402

# ---- P1 -------------------------------------------------------------------------------------------------------
_ref_get_combine_masks = ref_ve.get_combine_masks
ref_ve.get_combine_masks = lambda boxes, grid_size=7: _ref_get_combine_masks(boxes, grid_size).squeeze(1)


# ---- P2 -------------------------------------------------------------------------------------------------------
def _dual_embedding_forward(self, region_features, region_boxes, grid_features, grid_boxes):
    region_masks = ref_ve.generate_padding_mask(region_features, padding_idx=0)
    grid_masks = ref_ve.generate_padding_mask(grid_features, padding_idx=0)
    n, g2 = region_features.shape[1], grid_features.shape[1]
    region2grid = ref_ve.get_combine_masks(region_boxes, int(g2 ** 0.5))
    grid2region = region2grid.permute(0, 1, 3, 2)
    region2all = torch.cat([region_masks.expand(-1, -1, n, -1), region2grid], dim=-1)
    grid2all = torch.cat([grid2region, grid_masks.expand(-1, -1, g2, -1)], dim=-1)
    return ((self.region_proj(region_features), region_masks), (self.grid_proj(grid_features), grid_masks),
            (region2all, grid2all))


ref_ve.GeometricDualFeatureEmbedding.forward = _dual_embedding_forward

# ---- P3 -------------------------------------------------------------------------------------------------------
_ref_layer_forward = RefEncoderLayer.forward
_ROW_MASKS = {}   # number of queries -> that stream's (B,1,1,nq) padding mask, set per batch by DLCTReference


def _layer_forward(self, queries, keys, values, padding_mask, attention_mask, **kwargs):
    if padding_mask.shape[2] != 1:   # an attention mask was handed in as the row mask
        padding_mask = _ROW_MASKS[queries.shape[1]]
    return _ref_layer_forward(self, queries=queries, keys=keys, values=values, padding_mask=padding_mask,
                              attention_mask=attention_mask, **kwargs)


RefEncoderLayer.forward = _layer_forward


class DLCTReference(RefBaseTransformer):
    """The wiring the reference lacks, in the shape of its other architectures (models/standard_stransformer.py)."""

    def __init__(self, config, vocab):
        super().__init__(vocab)
        self.device = torch.device("cpu")
        self.vision_embedding = ref_build_vision_embedding(config.VISION_EMBEDDING)
        self.encoder = ref_build_encoder(config.ENCODER)
        self.decoder = ref_build_decoder(config.DECODER, vocab)

    def encoder_forward(self, f):
        (region, r_mask), (grid, g_mask), (region2all, grid2all) = self.vision_embedding(
            f.region_features, f.region_boxes, f.grid_features, f.grid_boxes)
        _ROW_MASKS.clear()
        _ROW_MASKS[region.shape[1]] = r_mask
        _ROW_MASKS[grid.shape[1]] = g_mask
        assert region.shape[1] != grid.shape[1]
        return self.encoder(region, f.region_boxes, r_mask, region2all, grid, f.grid_boxes, g_mask, grid2all)

    def forward(self, f):
        enc, mask = self.encoder_forward(f)
        return self.decoder(caption_tokens=f.caption_tokens, encoder_features=enc, encoder_attention_mask=mask)


def main():
    name, case = "dlct", CASES["dlct"]
    cfg_path = REPO / "openviic_b200" / "configs" / case["config"]
    ref_cfg = apply_overrides(ref_get_config(str(cfg_path)), case)
    vocab = synthetic.SyntheticVocab(case["vocab"], case["max_len"])
    torch.manual_seed(0)
    model = DLCTReference(ref_cfg.MODEL, vocab).eval()
    weights = synthetic.load_synthetic_weights(model, case["seed"])
    inputs = synthetic.synth_dual_inputs(ref_cfg.MODEL, case["batch"], case["n"], case["grid"], case["seed"])
    items = RefInstanceList()
    for key, value in inputs.items():
        items.set(key, value)
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

    our_cfg = apply_overrides(get_config(cfg_path), case)
    feats = (inputs["region_features"], inputs["grid_features"])
    boxes = (inputs["region_boxes"], inputs["grid_boxes"])
    logits_trace = []
    o_ids, o_logp = oracle.caption_beam_search(weights, our_cfg.MODEL, vocab, feats, boxes, beam=beam, out_size=1,
                                               logits_trace=logits_trace)
    o_ids_all, o_logp_all = oracle.caption_beam_search(weights, our_cfg.MODEL, vocab, feats, boxes, beam=beam, out_size=beam)
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
    print(f"[{name}] reference composition {ref_seconds:.2f}s  oracle-vs-reference {checks}  {'OK' if ok else 'MISMATCH'}")
    if not ok:
        raise SystemExit("oracle does not reproduce the patched reference composition")
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
    out = REPO / "tests" / "golden" / f"{name}.npz"
    np.savez_compressed(out, **fixture)
    print(f"[{name}] wrote {out.stat().st_size / 1024:.0f} KiB, captions ending in <eos>: {int(fixture['n_eos'])}/{b}, "
          f"masked region rows {int(enc_mask[..., :case['n']].sum())}")


if __name__ == "__main__":
    main()
