"""Deterministic synthetic vocabulary, weights and visual features.

There is no network for datasets or checkpoints, and the reference itself does not exist on the GPU
box, so every run (tests, golden fixtures, bench, smoke) uses inputs that any side can regenerate
from (parameter name, shape, seed) alone:

* weights: per-tensor ``numpy.random.default_rng([seed, crc32(name)])`` streams, scaled like the
  reference's initialisers (Xavier-sized projections, unit LayerNorm gains with jitter, N(0,1)
  word embeddings, ``m_k ~ N(0, 1/d_k)``, ``m_v ~ N(0, 1/m)``; attentions.py:34-42,151-152), and the
  vocabulary projection sharpened so that beams are separated like a trained model's;
* all values are rounded to bf16-representable numbers, so the fp32 oracle and the bf16 CUDA path
  consume bit-identical parameters and inputs (what remains is activation rounding only);
* the frozen ``decoder.pos_emb.weight`` sinusoid table is left as the module built it.

The same function seeds the *reference* model in ``oracle/ref_harness/gen_golden.py``.
"""

from __future__ import annotations

import zlib
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch


class SyntheticVocab:
    """The only vocabulary surface the model touches (base_transformer.py:13-14,33;
    decoders.py:82-83,90; text_embeddings.py:12,15)."""

    def __init__(self, size: int, max_caption_length: int):
        self.itos = ["<pad>", "<bos>", "<eos>", "<unk>"] + [f"w{i}" for i in range(4, size)]
        self.stoi = {w: i for i, w in enumerate(self.itos)}
        self.padding_idx, self.bos_idx, self.eos_idx, self.unk_idx = 0, 1, 2, 3
        self.max_caption_length = max_caption_length

    def __len__(self) -> int:
        return len(self.itos)

    def decode_caption(self, ids: torch.Tensor, join_words: bool = True):
        """ids (B,T) -> captions, stopping at <eos> (reference: data_utils/vocab.py:104-122)."""
        captions = []
        for row in ids.tolist():
            words = []
            for idx in row:
                if idx == self.eos_idx:
                    break
                if idx not in (self.padding_idx, self.bos_idx):
                    words.append(self.itos[idx])
            captions.append(" ".join(words) if join_words else words)
        return captions


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int, eos_idx: int = 2, pad_idx: int = 0) -> torch.Tensor:
    rng = _rng(seed, name)
    z = torch.from_numpy(rng.standard_normal(size=shape, dtype=np.float32))
    leaf = name.rsplit(".", 1)[-1]
    if name.endswith("layer_norm.weight"):
        t = 1.0 + 0.1 * z
    elif name.endswith("layer_norm.bias"):
        t = 0.05 * z
    elif leaf == "m_k":
        t = z / 64.0
    elif leaf == "m_v":
        t = z / 40.0
    elif "fc_gs." in name:
        t = 0.5 * z if leaf == "weight" else 0.1 + 0.05 * z
    elif name.endswith("word_emb.components.weight"):
        t = z.clone()
        t[pad_idx] = 0
    elif name == "decoder.fc.weight":
        t = 3.0 * z / shape[1] ** 0.5
    elif leaf == "bias":
        t = 0.02 * z
    elif len(shape) == 2:
        t = z / shape[1] ** 0.5
    else:
        t = 0.02 * z
    return bf16_round(t)


SKIP = ("decoder.pos_emb.weight",)
EOS_OFFSET = 4.5


def synth_state_dict(named_shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int) -> Dict[str, torch.Tensor]:
    """Weights for every (name, shape) except the frozen position table.

    The <eos> row of the vocabulary projection gets a component along the last decoder LayerNorm's
    bias (a constant part of every decoder output), i.e. a constant logit offset of EOS_OFFSET, so that
    <eos> is competitive at every step and the finished-beam path (seq_mask, -999 sentinel, <pad>
    feeding, padded-row zeroing) is exercised by random-weight models.
    """
    sd = {name: synth_tensor(name, tuple(shape), seed) for name, shape in named_shapes if name not in SKIP}
    last_ln = sorted((k for k in sd if k.startswith("decoder.layers.") and k.endswith(".pwff.layer_norm.bias")),
                     key=lambda k: int(k.split(".")[2]))
    if last_ln and "decoder.fc.weight" in sd:
        beta = sd[last_ln[-1]]
        fc = sd["decoder.fc.weight"]
        fc[2] = bf16_round(fc[2] + EOS_OFFSET * beta / float(beta.pow(2).sum()))
    return sd


def load_synthetic_weights(model: torch.nn.Module, seed: int) -> Dict[str, torch.Tensor]:
    """Overwrite ``model``'s parameters in place (works for the reference model and for ours)."""
    current = model.state_dict()
    sd = synth_state_dict(((k, tuple(v.shape)) for k, v in current.items() if v.dtype.is_floating_point), seed)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items() if v.dtype.is_floating_point}


def synth_features(batch: int, n: int, d_feature: int, seed: int, ragged: bool) -> torch.Tensor:
    """(B,n,D) fp32 bf16-representable features; ``ragged`` zeroes a random suffix of rows per image
    (valid count ~ U[n/2, n]) so the padding-mask path (models/utils.py:60) is exercised."""
    rng = _rng(seed, f"features/{batch}/{n}/{d_feature}")
    x = torch.from_numpy(rng.standard_normal(size=(batch, n, d_feature), dtype=np.float32))
    if ragged:
        valid = rng.integers(low=max(1, n // 2), high=n + 1, size=batch)
        valid[0] = n
        for b, k in enumerate(valid):
            x[b, int(k):] = 0
    return bf16_round(x)


def synth_boxes(batch: int, n: int, seed: int) -> torch.Tensor:
    """(B,n,4) boxes (x1,y1,x2,y2): xy ~ U[0,.5], wh ~ U[.05,.55]."""
    rng = _rng(seed, f"boxes/{batch}/{n}")
    xy = rng.uniform(0.0, 0.5, size=(batch, n, 2)).astype(np.float32)
    wh = rng.uniform(0.05, 0.55, size=(batch, n, 2)).astype(np.float32)
    return torch.from_numpy(np.concatenate([xy, xy + wh], axis=-1))


def synth_grid_boxes(batch: int, grid_size: int) -> torch.Tensor:
    """(B, g*g, 4) boxes of the g x g grid cells, row-major, in [0,1] image coordinates."""
    k = torch.arange(grid_size * grid_size)
    x, y = (k % grid_size).float(), torch.div(k, grid_size, rounding_mode="floor").float()
    cells = torch.stack([x / grid_size, y / grid_size, (x + 1) / grid_size, (y + 1) / grid_size], dim=-1)
    return cells.unsqueeze(0).expand(batch, -1, -1).contiguous()


def synth_dual_inputs(model_cfg, batch: int, n_regions: int, grid_size: int, seed: int) -> Dict[str, torch.Tensor]:
    """Inputs of the dual-path (region + grid) models: ragged region features with boxes, a full grid with its cells."""
    ve = model_cfg.VISION_EMBEDDING
    return {
        "region_features": synth_features(batch, n_regions, ve.D_REGION_FEATURE, seed, ragged=True),
        "region_boxes": synth_boxes(batch, n_regions, seed),
        "grid_features": synth_features(batch, grid_size * grid_size, ve.D_GRID_FEATURE, seed + 1, ragged=False),
        "grid_boxes": synth_grid_boxes(batch, grid_size),
    }


def synth_adaptive_inputs(case):
    """(queries, keys, language signals, key mask (B,1,1,nk)) for the operator-level adaptive-attention case."""
    b, nq, nk, d = case["batch"], case["nq"], case["nk"], case["config"]["D_MODEL"]
    rng = _rng(case["seed"], "adaptive_attention")
    q, k, s = (bf16_round(torch.from_numpy(rng.standard_normal(size=shape, dtype=np.float32)))
               for shape in ((b, nq, d), (b, nk, d), (b, nq, d)))
    mask = torch.zeros(b, 1, 1, nk, dtype=torch.bool)
    mask[1, ..., nk // 2:] = True
    mask[2, ..., ::3] = True
    return q, k, s, mask


def feature_field(model_cfg) -> str:
    return "grid_features" if model_cfg.ARCHITECTURE == "StandardTransformerUsingGrid" else "region_features"


def needs_boxes(model_cfg) -> bool:
    return model_cfg.ENCODER.ARCHITECTURE == "GeometricEncoder"


def synth_inputs(model_cfg, batch: int, n: int, seed: int, ragged: Optional[bool] = None):
    """Feature tensor (+ boxes) for a model config; grid inputs are never ragged."""
    field = feature_field(model_cfg)
    if ragged is None:
        ragged = field == "region_features"
    feats = synth_features(batch, n, model_cfg.VISION_EMBEDDING.D_FEATURE, seed, ragged)
    boxes = synth_boxes(batch, n, seed) if needs_boxes(model_cfg) else None
    return field, feats, boxes


def synth_captions(batch: int, max_len: int, vocab_size: int, seed: int, bos_idx: int = 1, eos_idx: int = 2,
                   pad_idx: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Training captions as FeatureDataset.__getitem__ builds them (data_utils/dataset.py:56-61 over
    Vocab.encode_caption, data_utils/vocab.py:97-102): the encoded caption is <bos> w.. <eos> <pad>.. in max_len tokens;
    the targets (`shifted_right_caption_tokens`) are that row shifted left by one, and the INPUT row has its <eos>
    replaced by <pad>.  3 .. max_len - 2 words per caption.  -> (tokens, targets), int64 (batch, max_len)."""
    rng = _rng(seed, "captions")
    tokens = np.full((batch, max_len), pad_idx, dtype=np.int64)
    for i in range(batch):
        words = int(rng.integers(3, max(4, max_len - 1)))
        words = min(words, max_len - 2)
        tokens[i, 0] = bos_idx
        tokens[i, 1:1 + words] = rng.integers(4, vocab_size, size=words)
        tokens[i, 1 + words] = eos_idx
    targets = np.full_like(tokens, pad_idx)
    targets[:, :-1] = tokens[:, 1:]
    tokens[tokens == eos_idx] = pad_idx
    return torch.from_numpy(tokens), torch.from_numpy(targets)


def synth_train_batches(model_cfg, case: dict):
    """The batches of a training case (oracle/cases.py TRAIN_CASES): [(field, feats, tokens, targets, boxes)] * steps."""
    out = []
    for i in range(case["steps"]):
        field, feats, boxes = synth_inputs(model_cfg, case["batch"], case["n"], case["seed"] + 100 * i)
        tokens, targets = synth_captions(case["batch"], case["max_len"], case["vocab"], case["seed"] + 100 * i)
        out.append((field, bf16_round(feats), tokens, targets, boxes))
    return out


def synth_rewards(batch: int, beam: int, seed: int) -> torch.Tensor:
    """Stand-in CIDEr rewards (batch, beam) fp32 in [0, 2) for the self-critical step's parity cases."""
    return torch.from_numpy(_rng(seed, "rewards").random((batch, beam)).astype(np.float32) * 2.0)


def boost_eos(model: torch.nn.Module, weights: Dict[str, torch.Tensor], eos_idx: int, scale: float) -> None:
    """Scales the <eos> row of the vocabulary projection (in the model and in the weight dict): synthetic weights almost
    never emit <eos>; with the row scaled a good share of the beams finish early, which is what the parity cases of the
    self-critical step need (finished beams, <pad> after <eos>)."""
    with torch.no_grad():
        row = bf16_round(weights["decoder.fc.weight"][eos_idx] * scale)
        weights["decoder.fc.weight"][eos_idx] = row
        model.state_dict()["decoder.fc.weight"][eos_idx] = row
