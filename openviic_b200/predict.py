"""The evaluation loop's per-batch body, host rows in -> caption text out
(reference: trainers/vi_trainer.py:242-251 -- ``items.to(device)``, ``model.beam_search(items, batch_size,
beam_size, out_size=1)``, ``vocab.decode_caption(outs.view(-1, T), join_words=False)``, groupby collapse).

    per-image (n_i, D) fp32 rows --FeatureBatcher--> pinned bf16 (B, n, D) --CaptionEngine.caption_host-->
    pinned int64 (B, 1, T) ids --Vocab.decode_predictions--> B strings

Every stage is native (host_glue.cpp, the CUDA engine); Python only hands pointers over.  ``submit`` / ``collect``
split the call so that a caller keeps several batches in flight (one predictor per stream), the way bench.py's
e2e loop does.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .data_utils import FeatureBatcher


class CaptionPredictor:
    def __init__(self, model, vocab, max_batch: int, max_rows: int, beam_size: int = 5, threads: Optional[int] = None):
        """``model``: a built architecture (``build_model``) on a CUDA device; ``vocab``: a ``data_utils.Vocab``
        (anything with ``decode_predictions``).  Batches are padded to ``max_rows`` rows per image: one shape,
        one captured CUDA graph; all-zero rows are padding to the model, as in the reference (models/utils.py:60)."""
        from .synthetic import needs_boxes
        self.model, self.vocab = model, vocab
        self.max_batch, self.max_rows, self.beam_size = max_batch, max_rows, beam_size
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("CaptionPredictor needs the model on a CUDA device (the caption path has no CPU fallback)")
        cfg = model.model_config
        self.with_boxes = needs_boxes(cfg)
        self.batcher = FeatureBatcher(max_batch, max_rows, cfg.VISION_EMBEDDING.D_FEATURE,
                                      box_width=4 if self.with_boxes else 0, slots=2, threads=threads)
        self.engine = model.engine(max_batch, max_rows, beam_size)
        T = self.engine.max_len
        self._ids = torch.empty((max_batch, 1, T), dtype=torch.int64).pin_memory()
        self._logp = torch.empty((max_batch, 1, T), dtype=torch.float32).pin_memory()
        self._pending = 0

    def submit(self, features: Sequence, boxes: Optional[Sequence] = None) -> None:
        """Collate on the host and enqueue H2D + encoder + beam search + D2H on the current stream; returns at once."""
        if self._pending:
            raise RuntimeError("collect() the previous batch first (one batch in flight per predictor)")
        if self.with_boxes and boxes is None:
            raise ValueError("this architecture reads region boxes")
        feats, bx = self.batcher.collate(features, boxes if self.with_boxes else None, pad_to=self.max_rows)
        b = feats.shape[0]
        self.engine.caption_host(feats, bx, out_size=1, use_graph=True,
                                 out=(self._ids[:b], self._logp[:b]), sync=False)
        self._pending = b

    def collect(self) -> List[str]:
        """Wait for the submitted batch and turn its ids into text."""
        if not self._pending:
            raise RuntimeError("nothing submitted")
        torch.cuda.current_stream(self.device).synchronize()
        b, self._pending = self._pending, 0
        return self.vocab.decode_predictions(self._ids[:b, 0])

    def predict(self, features: Sequence, boxes: Optional[Sequence] = None) -> List[str]:
        self.submit(features, boxes)
        return self.collect()

    def last_log_probs(self, batch: int) -> torch.Tensor:
        """(batch, T) per-token log-probs of the last collected batch's captions."""
        return self._logp[:batch, 0].clone()


def _batches(dataset, batch_size: int):
    for start in range(0, len(dataset), batch_size):
        yield [dataset[i] for i in range(start, min(start + batch_size, len(dataset)))]


def get_predictions(predictor, dataset, batch_size: int, get_scores: bool = True) -> dict:
    """The test-set loop of the reference (trainers/vi_trainer.py:229-276) over a ``DictionaryDataset``: per batch
    ``{"image_id", "filename", "gens", "gts"}`` with keys ``"<batch>_<index>"``, plus the corpus scores.

    Feature loading of batch k+1 runs while batch k is on the GPU (``submit`` returns at once).  Differences, on
    purpose: ``image_id`` holds the dataset's ids (the reference's ``items.image_id`` is always ``None`` because its
    samples carry no such field), and of the reference's four metrics only CIDEr is computed (the others shell out to
    Java tools).  ``predictor`` needs ``submit(features, boxes)`` / ``collect()`` (``CaptionPredictor``)."""
    from .evaluation import Cider
    results, overall_gens, overall_gts = [], {}, {}
    pending = None

    def finish(entry):
        it, samples, ids = entry
        texts = predictor.collect()
        gens = {"%d_%d" % (it, i): text for i, text in enumerate(texts)}
        gts = {"%d_%d" % (it, i): s.captions for i, s in enumerate(samples)}
        overall_gens.update({k: [v] for k, v in gens.items()})
        overall_gts.update(gts)
        results.append({"image_id": ids, "filename": [s.filename for s in samples], "gens": gens, "gts": gts})

    for it, samples in enumerate(_batches(dataset, batch_size)):      # loading batch `it` overlaps batch `it - 1`
        field = "region_features" if "region_features" in samples[0] else "grid_features"
        feats = [s[field] for s in samples]
        boxes = [s["region_boxes"] for s in samples] if "region_boxes" in samples[0] else None
        if pending is not None:
            finish(pending)
        predictor.submit(feats, boxes)
        start = it * batch_size
        pending = (it, samples, list(dataset.image_ids[start:start + len(samples)]))
    if pending is not None:
        finish(pending)
    scores = {}
    if get_scores and overall_gts:
        scores["CIDEr"] = float(Cider().compute_score(overall_gts, overall_gens)[0])
    return {"results": results, **scores}


def evaluate_metrics(predictor, dataset, batch_size: int) -> dict:
    """The validation loop of the reference (trainers/vi_trainer.py:78-98): corpus scores only."""
    out = get_predictions(predictor, dataset, batch_size, get_scores=True)
    return {k: v for k, v in out.items() if k != "results"}
