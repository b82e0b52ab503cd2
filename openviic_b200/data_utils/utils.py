"""Caption pre-processing and the feature collate (reference: data_utils/utils.py:6-83,126-127 and
utils/instance.py:32-55,156-171)."""

from __future__ import annotations

import ctypes as C
import os
import re
from typing import Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import cabi
from ..utils.instance import Instance, InstanceList

_QUOTES = re.compile(r"[“”]")
_SPACED = re.compile(r"([!?:;,\"'()\[\]/.$&*])")


def get_tokenizer(tokenizer) -> Callable[[str], str]:
    """None -> identity, a callable -> itself, a name -> that word segmenter (reference data_utils/utils.py:6-55).
    The named segmenters are third-party packages; a missing one raises ImportError."""
    if tokenizer is None:
        return lambda s: s
    if callable(tokenizer):
        return tokenizer
    if tokenizer == "pyvi":
        from pyvi import ViTokenizer
        return ViTokenizer.tokenize
    if tokenizer == "spacy":
        from spacy.lang.vi import Vietnamese
        return Vietnamese()
    if tokenizer == "vncorenlp":
        from vncorenlp import VnCoreNLP
        annotator = VnCoreNLP(address="http://127.0.0.1", port=9000)
        return lambda s: " ".join(annotator.tokenize(s)[0])
    raise ValueError(f"unknown tokenizer {tokenizer!r}")


def preprocess_caption(caption: str, tokenizer=None) -> List[str]:
    """Caption text -> lower-cased words with every punctuation mark of the reference's list split off
    (reference data_utils/utils.py:57-80: curly quotes straightened, then ``! ? : ; , " ' ( [ ) ] / . $ & *``
    each surrounded by spaces, lower-casing, the word segmenter, whitespace split)."""
    caption = _SPACED.sub(r" \1 ", _QUOTES.sub('"', caption))
    caption = get_tokenizer(tokenizer)(caption.lower())
    return caption.strip().split()


def collate_fn(samples: List[Instance]) -> InstanceList:
    """The DataLoader collate of the reference (data_utils/utils.py:126-127)."""
    return InstanceList(samples)


def _as_rows(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim != 2:
        raise ValueError(f"per-image features must be (rows, width), got shape {tuple(x.shape)}")
    return x


class FeatureBatcher:
    """Per-image (n_i, D) fp32 feature rows -> one zero-padded (B, n, D) batch in pinned host memory, as bf16
    (what ``CaptionEngine.caption_host`` copies to the GPU), in one native multi-threaded pass
    (``cap_host_collate_bf16``).

    The reference does this in three steps: ``InstanceList.__init__`` pads every image with a fresh ``torch.zeros``
    block and concatenates (utils/instance.py:42-49,156-171), the DataLoader hands the fp32 batch over, and
    ``items.to(device)`` stages it from pageable memory (trainers/vi_trainer.py:244).  ``n`` is the longest image of
    the batch as there, or ``pad_to`` when given (a fixed shape keeps one CUDA graph).  Boxes, when present, stay fp32.

    ``slots`` staging buffers rotate so that batch k+1 can be collated while batch k's H2D copy is still in flight.
    """

    def __init__(self, max_batch: int, max_rows: int, width: int, box_width: int = 0, slots: int = 2,
                 threads: Optional[int] = None, pin: Optional[bool] = None):
        self.max_batch, self.max_rows, self.width, self.box_width = max_batch, max_rows, width, box_width
        self.threads = threads or min(16, os.cpu_count() or 1)
        pin = torch.cuda.is_available() if pin is None else pin

        def staging(shape, dtype):
            t = torch.empty(shape, dtype=dtype)
            return t.pin_memory() if pin else t

        self._feats = [staging((max_batch, max_rows, width), torch.bfloat16) for _ in range(slots)]
        self._boxes = [staging((max_batch, max_rows, box_width), torch.float32) for _ in range(slots)] if box_width else None
        self._next = 0

    def _pointers(self, rows: Sequence[np.ndarray], width: int):
        counts = np.empty(len(rows), dtype=np.int32)
        ptrs = (C.c_void_p * len(rows))()
        for i, r in enumerate(rows):
            if r.shape[1] != width:
                raise ValueError(f"image {i}: rows are {r.shape[1]} wide, the batcher was built for {width}")
            counts[i] = r.shape[0]
            ptrs[i] = r.ctypes.data
        return ptrs, counts

    def collate(self, features: Sequence[Union[np.ndarray, torch.Tensor]],
                boxes: Optional[Sequence[Union[np.ndarray, torch.Tensor]]] = None,
                pad_to: Optional[int] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        B = len(features)
        if B == 0 or B > self.max_batch:
            raise ValueError(f"batch of {B} images, the batcher holds 1..{self.max_batch}")
        rows = [_as_rows(f) for f in features]
        n = max(r.shape[0] for r in rows) if pad_to is None else pad_to
        if n > self.max_rows:
            raise ValueError(f"{n} rows per image, the batcher holds {self.max_rows}")
        slot, self._next = self._next, (self._next + 1) % len(self._feats)
        ptrs, counts = self._pointers(rows, self.width)
        out = self._feats[slot].view(-1)[: B * n * self.width].view(B, n, self.width)
        cabi.call("cap_host_collate_bf16", ptrs, counts.ctypes.data, B, n, self.width, out.data_ptr(), self.threads)
        out_boxes = None
        if boxes is not None:
            if not self.box_width:
                raise ValueError("this batcher was built without boxes")
            brows = [_as_rows(b) for b in boxes]
            if [b.shape[0] for b in brows] != [r.shape[0] for r in rows]:
                raise ValueError("boxes and features disagree on the number of rows per image")
            bptrs, bcounts = self._pointers(brows, self.box_width)
            out_boxes = self._boxes[slot].view(-1)[: B * n * self.box_width].view(B, n, self.box_width)
            cabi.call("cap_host_collate_f32", bptrs, bcounts.ctypes.data, B, n, self.box_width, out_boxes.data_ptr(),
                      self.threads)
        return out, out_boxes
