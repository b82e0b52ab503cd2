"""Vocabulary: captions <-> id vectors (reference: data_utils/vocab.py:12-122).

Same constructor contract (the config node with ``VOCAB`` / ``JSON_PATH`` / ``MIN_FREQ``), same attributes
(``itos``, ``stoi``, ``freqs``, ``specials``, ``padding_idx`` ..., ``max_caption_length``), same ordering rule
(specials first, then by falling frequency, ties alphabetically).  ``decode_caption`` runs natively
(``cap_vocab_decode`` in csrc/host_glue.cu): at the GPU path's caption rate the reference's per-token Python
loop is slower than the captions arrive.

Not carried over (they need downloads that do not exist here): ``PRETRAINED_LANGUAGE_MODEL`` /
``USE_MAPPING`` (HF tokenizers) and ``WORD_EMBEDDING`` (pretrained vectors); selecting them raises.
"""

from __future__ import annotations

import ctypes as C
import json
from collections import Counter
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from .. import cabi
from .utils import preprocess_caption


class Vocab:
    def __init__(self, config):
        self.tokenizer = config.VOCAB.TOKENIZER
        if config.VOCAB.PRETRAINED_LANGUAGE_MODEL is not None or config.VOCAB.USE_MAPPING:
            raise NotImplementedError("vocabularies tied to a pretrained language model need its tokenizer files "
                                      "(reference data_utils/vocab.py:20-25,68-77); not available offline")
        if config.VOCAB.WORD_EMBEDDING is not None:
            raise NotImplementedError("pretrained word vectors (reference data_utils/vocab.py:81-83) need downloads")
        self.padding_token = config.VOCAB.PAD_TOKEN
        self.bos_token = config.VOCAB.BOS_TOKEN
        self.eos_token = config.VOCAB.EOS_TOKEN
        self.unk_token = config.VOCAB.UNK_TOKEN
        self.make_vocab([config.JSON_PATH.TRAIN, config.JSON_PATH.DEV, config.JSON_PATH.TEST])
        self._index(self.freqs, max(config.MIN_FREQ, 1))

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_itos(cls, itos: Sequence[str], max_caption_length: int,
                  specials: Sequence[str] = ("<pad>", "<bos>", "<eos>", "<unk>")) -> "Vocab":
        """A vocabulary from an existing word list (a saved ``itos``); the four specials must be in it."""
        self = cls.__new__(cls)
        self.tokenizer = None
        self.padding_token, self.bos_token, self.eos_token, self.unk_token = specials
        self.freqs = Counter()
        self.output_cats = set()
        self.max_caption_length = int(max_caption_length)
        self._finish(list(itos))
        return self

    def make_vocab(self, json_dirs) -> None:
        """Word frequencies and the longest caption (+2 for bos/eos) over the annotation files
        (reference data_utils/vocab.py:84-94)."""
        self.freqs = Counter()
        self.output_cats = set()
        self.max_caption_length = 0
        for json_dir in json_dirs:
            with open(json_dir, encoding="utf-8") as fh:
                annotations = json.load(fh)["annotations"]
            for ann in annotations:
                words = preprocess_caption(ann["caption"], self.tokenizer)
                self.freqs.update(words)
                self.max_caption_length = max(self.max_caption_length, len(words) + 2)

    def _index(self, freqs: Counter, min_freq: int) -> None:
        specials = [self.padding_token, self.bos_token, self.eos_token, self.unk_token]
        counted = [(w, f) for w, f in freqs.items() if w not in specials and f >= min_freq]
        counted.sort(key=lambda wf: (-wf[1], wf[0]))   # falling frequency, ties alphabetically
        self._finish(specials + [w for w, _ in counted])

    def _finish(self, itos: List[str]) -> None:
        self.itos = itos
        self.stoi = {tok: i for i, tok in enumerate(itos)}
        self.padding_idx = self.stoi[self.padding_token]
        self.bos_idx = self.stoi[self.bos_token]
        self.eos_idx = self.stoi[self.eos_token]
        self.unk_idx = self.stoi[self.unk_token]
        self.specials = [self.padding_token, self.bos_token, self.eos_token, self.unk_token]
        self.mapping = None
        self.word_embeddings = None
        self._native = None
        # words that str.split() would cut or drop make "collapse over vocabulary entries" differ from the
        # reference's "collapse over the split words": those vocabularies take the collapse in Python
        self._plain_words = all(w and w.split() == [w] for w in itos)

    def __len__(self) -> int:
        return len(self.itos)

    def __eq__(self, other) -> bool:
        return (isinstance(other, Vocab) and self.freqs == other.freqs and self.stoi == other.stoi
                and self.itos == other.itos)

    def __del__(self):
        handle, self._native = getattr(self, "_native", None), None
        if handle is not None:
            try:
                cabi.call("cap_vocab_destroy", handle)
            except Exception:
                pass

    def extend(self, v: "Vocab", sort: bool = False) -> None:
        for w in (sorted(v.itos) if sort else v.itos):
            if w not in self.stoi:
                self.itos.append(w)
                self.stoi[w] = len(self.itos) - 1
        self._finish(self.itos)

    # ------------------------------------------------------------------ text -> ids
    def encode_caption(self, caption: List[str]) -> torch.Tensor:
        """Words -> (max_caption_length,) int64: bos, the words (unknown -> unk), eos, then padding
        (reference data_utils/vocab.py:96-102; a caption longer than max_caption_length - 2 raises IndexError)."""
        vec = torch.full((self.max_caption_length,), self.padding_idx, dtype=torch.long)
        ids = [self.bos_idx] + [self.stoi.get(tok, self.unk_idx) for tok in caption] + [self.eos_idx]
        if len(ids) > self.max_caption_length:
            raise IndexError(f"caption of {len(caption)} words does not fit max_caption_length {self.max_caption_length}")
        vec[: len(ids)] = torch.tensor(ids, dtype=torch.long)
        return vec

    # ------------------------------------------------------------------ ids -> text
    def _handle(self):
        if self._native is None:
            encoded = [w.encode("utf-8") for w in self.itos]
            offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
            np.cumsum([len(b) for b in encoded], out=offsets[1:])
            blob = b"".join(encoded)
            special = np.array([w in self.specials for w in self.itos], dtype=np.uint8)
            out = C.c_void_p()
            cabi.call("cap_vocab_create", C.c_char_p(blob), offsets.ctypes.data, len(encoded), special.ctypes.data,
                      int(self.eos_idx), C.byref(out))
            self._native = out
        return self._native

    def _decode(self, caption_vecs: Union[torch.Tensor, np.ndarray], collapse: bool) -> List[str]:
        if isinstance(caption_vecs, torch.Tensor):
            caption_vecs = caption_vecs.detach().cpu().numpy()
        ids = np.ascontiguousarray(caption_vecs, dtype=np.int64)
        if ids.ndim != 2:
            raise ValueError(f"caption_vecs must be (bs, max_length), got shape {tuple(ids.shape)}")
        n, T = ids.shape
        if n == 0:
            return []
        longest = int(np.diff(self._offsets()).max())
        buf = np.empty(n * (T * (longest + 1) + 1), dtype=np.uint8)
        used = C.c_int64(0)
        try:
            cabi.call("cap_vocab_decode", self._handle(), ids.ctypes.data, n, T, int(collapse), buf.ctypes.data,
                      buf.size, C.byref(used))
        except RuntimeError as err:
            if "outside the vocabulary" in str(err):
                raise IndexError(str(err)) from None   # the reference's itos[idx] raises IndexError
            raise
        text = buf[: used.value].tobytes().decode("utf-8")
        captions = text.split("\n")
        captions.pop()   # every caption is terminated, not separated, by the newline
        return captions

    def _offsets(self) -> np.ndarray:
        if getattr(self, "_word_offsets", None) is None or len(self._word_offsets) != len(self.itos) + 1:
            self._word_offsets = np.concatenate([[0], np.cumsum([len(w.encode("utf-8")) for w in self.itos])])
        return self._word_offsets

    def decode_caption(self, caption_vecs: torch.Tensor, join_words: bool = True,
                       collapse_repeats: bool = False):
        """(bs, max_length) ids -> captions: the non-special words up to the first eos
        (reference data_utils/vocab.py:104-122).  ``join_words=False`` returns word lists.

        ``collapse_repeats`` (extension) also applies the trainers' ``itertools.groupby`` pass, which keeps one
        of every run of equal consecutive words (reference trainers/vi_trainer.py:251)."""
        native_collapse = collapse_repeats and self._plain_words
        captions = self._decode(caption_vecs, native_collapse)
        if collapse_repeats and not native_collapse:
            captions = [" ".join(_collapse(c.strip().split())) for c in captions]
        if join_words:
            return captions
        return [c.strip().split() for c in captions]

    def decode_predictions(self, caption_vecs: torch.Tensor) -> List[str]:
        """What the evaluation loop makes of the beam search output: words decoded, consecutive duplicates
        collapsed, joined by spaces (reference trainers/vi_trainer.py:248-251)."""
        if self._plain_words:
            return self._decode(caption_vecs, True)
        return [" ".join(_collapse(words)) for words in self.decode_caption(caption_vecs, join_words=False)]


def _collapse(words: List[str]) -> List[str]:
    return [w for i, w in enumerate(words) if i == 0 or w != words[i - 1]]
