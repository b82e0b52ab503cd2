"""CIDEr-D with the reference's interface (evaluation/cider/cider.py:12-38), computed natively.

    Cider(gts)                      # document frequencies over a corpus: {key: [caption, ...]}
    Cider().compute_score(gts, res) # -> (mean, per-key scores); res = {key: [hypothesis]}

Captions are whitespace-separated words, as in the reference (``precook`` splits on whitespace,
evaluation/cider/cider_scorer.py:9-24).  The Python side only numbers the words; n-gram counting, tf-idf
weighting, the clipped cosine and the length penalty run in ``csrc/host_cider.cpp`` over all host threads.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .. import cabi


class Cider:
    def __init__(self, gts: Optional[Dict[str, Sequence[str]]] = None, n: int = 4, sigma: float = 6.0,
                 threads: Optional[int] = None):
        self._n, self._sigma = n, sigma
        self._threads = threads or min(16, os.cpu_count() or 1)
        self._word_ids: Dict[str, int] = {}
        self._cooked: Dict[str, np.ndarray] = {}     # caption string -> int32 word ids (references repeat every epoch)
        handle = C.c_void_p()
        cabi.call("cap_cider_create", n, float(sigma), C.byref(handle))
        self._h = handle
        self.ref_len = None
        if gts is not None:
            tokens, cap_off, img_off = self._ragged(list(gts.values()))
            n_images = len(img_off) - 1
            self.ref_len = np.log(float(n_images))
            table = self._log_table(n_images)
            cabi.call("cap_cider_set_corpus", self._h, tokens.ctypes.data, cap_off.ctypes.data, img_off.ctypes.data,
                      n_images, float(self.ref_len), table.ctypes.data, table.size)

    def __del__(self):
        handle, self._h = getattr(self, "_h", None), None
        if handle is not None:
            try:
                cabi.call("cap_cider_destroy", handle)
            except Exception:
                pass

    def __str__(self) -> str:
        return "CIDEr"

    # ------------------------------------------------------------------ strings -> ragged id arrays
    def _ids(self, caption: str) -> np.ndarray:
        ids = self._cooked.get(caption)
        if ids is None:
            table = self._word_ids
            ids = np.fromiter((table.setdefault(w, len(table)) for w in caption.split()), dtype=np.int32)
            self._cooked[caption] = ids
        return ids

    def _ragged(self, groups: Sequence[Sequence[str]]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Groups of captions -> (tokens, caption offsets, group offsets)."""
        arrays: List[np.ndarray] = []
        group_off = np.zeros(len(groups) + 1, dtype=np.int64)
        for g, captions in enumerate(groups):
            if isinstance(captions, str):
                raise TypeError("every entry must be a list of captions, not a single string")
            arrays.extend(self._ids(c) for c in captions)
            group_off[g + 1] = len(arrays)
        cap_off = np.zeros(len(arrays) + 1, dtype=np.int64)
        if arrays:
            np.cumsum([a.size for a in arrays], out=cap_off[1:])
        tokens = np.concatenate(arrays) if arrays else np.zeros(0, dtype=np.int32)
        return np.ascontiguousarray(tokens, dtype=np.int32), cap_off, group_off

    @staticmethod
    def _log_table(largest: int) -> np.ndarray:
        table = np.zeros(largest + 1, dtype=np.float64)
        table[1:] = np.log(np.arange(1, largest + 1, dtype=np.float64))
        return table

    # ------------------------------------------------------------------ scoring
    def compute_score(self, gts: Dict[str, Sequence[str]], res: Dict[str, Sequence[str]]):
        """``gts[key]``: the reference captions of an image, ``res[key]``: a one-element list holding the hypothesis.
        Returns (corpus mean, per-key scores in the order of ``gts``) like evaluation/cider/cider.py:28-38."""
        assert gts.keys() == res.keys()
        keys = list(gts.keys())
        if not keys:
            return float("nan"), np.zeros(0)
        hyp_tokens, hyp_off, _ = self._ragged([[res[k][0]] for k in keys])
        ref_tokens, ref_cap_off, ref_group_off = self._ragged([gts[k] for k in keys])
        scores = np.empty(len(keys), dtype=np.float64)
        table = self._log_table(len(keys)) if self.ref_len is None else np.zeros(0)
        cabi.call("cap_cider_score", self._h, hyp_tokens.ctypes.data, hyp_off.ctypes.data, len(keys),
                  ref_tokens.ctypes.data, ref_cap_off.ctypes.data, ref_group_off.ctypes.data,
                  float(np.log(float(len(keys)))), table.ctypes.data if table.size else None, table.size,
                  scores.ctypes.data, self._threads)
        if len(self._cooked) > 2_000_000:   # hypotheses are new strings every step: do not grow without bound
            self._cooked.clear()
        return np.mean(scores), scores
