"""Annotation + pre-computed feature datasets feeding the caption path
(reference: data_utils/dataset.py:12-72 ``FeatureDataset``, :74-132 ``DictionaryDataset``).

On disk, as the reference expects it: one COCO-style JSON (``images`` with ``id`` / ``file_name``, ``annotations``
with ``image_id`` / ``caption``) and one ``{image_id}.npy`` per image under ``config.FEATURE_PATH.FEATURES`` holding a
pickled dict of arrays (``region_features`` (n, D), ``region_boxes`` (n, 4), ``grid_features`` ...).  Samples are
``Instance`` objects with the reference's field names and order; batches are built by ``data_utils.collate_fn`` (the
reference's) or, on the fast path, by ``FeatureBatcher`` from the per-image arrays.
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List

import numpy as np
import torch
from torch.utils.data import Dataset

from ..utils.instance import Instance
from .utils import preprocess_caption


class _AnnotatedFeatures(Dataset):
    """What both datasets share: the parsed annotation file, the vocabulary and the per-image feature files."""

    def __init__(self, json_path: str, vocab, config) -> None:
        super().__init__()
        self.vocab = vocab
        self.image_features_path = config.FEATURE_PATH.FEATURES
        with open(json_path, encoding="utf-8") as fh:
            self._index(json.load(fh))

    def _index(self, json_data: Dict) -> None:
        raise NotImplementedError

    def _words(self, caption: str) -> List[str]:
        return preprocess_caption(caption, self.vocab.tokenizer)

    def load_features(self, image_id) -> Dict[str, Any]:
        """The pickled dict of arrays of one image (``np.save(path, {...})``; ``[()]`` unwraps the 0-d object array)."""
        return np.load(Path(self.image_features_path) / f"{image_id}.npy", allow_pickle=True)[()]


class FeatureDataset(_AnnotatedFeatures):
    """One sample per ANNOTATION: teacher-forcing tokens + the image's features (XE training / validation loss)."""

    def _index(self, json_data: Dict) -> None:
        self.annotations = self.load_json(json_data)

    def load_json(self, json_data: Dict) -> List[Dict]:
        # the reference scans the image list once per annotation (data_utils/dataset.py:29-42); one id -> file name
        # table gives the same records.  An annotation whose image is missing raises KeyError here -- the reference
        # silently repeats the previous record.
        file_of = {}
        for image in json_data["images"]:
            file_of.setdefault(image["id"], image["file_name"])
        return [dict(caption=self._words(a["caption"]), image_id=a["image_id"], filename=file_of[a["image_id"]])
                for a in json_data["annotations"]]

    @property
    def captions(self) -> List[List[str]]:
        return [record["caption"] for record in self.annotations]

    def __len__(self) -> int:
        return len(self.annotations)

    def __getitem__(self, idx: int) -> Instance:
        record = self.annotations[idx]
        pad, eos = self.vocab.padding_idx, self.vocab.eos_idx
        tokens = self.vocab.encode_caption(record["caption"])          # bos w1 .. wk eos pad ...
        targets = torch.cat([tokens[1:], tokens.new_full((1,), pad)])   # w1 .. wk eos pad ... pad
        inputs = tokens.masked_fill(tokens == eos, pad)                 # the decoder input never holds eos
        return Instance(caption_tokens=inputs, shifted_right_caption_tokens=targets,
                        **self.load_features(record["image_id"]))


class DictionaryDataset(_AnnotatedFeatures):
    """One sample per IMAGE with all of its reference captions (evaluation, self-critical training)."""

    def _index(self, json_data: Dict) -> None:
        self.image_ids, self.filenames, self.captions_with_image = self.load_json(json_data)

    def load_json(self, json_data: Dict):
        by_image: Dict[Any, List[str]] = {image["id"]: [] for image in json_data["images"]}
        file_of = {image["id"]: image["file_name"] for image in json_data["images"]}
        for ann in json_data["annotations"]:
            by_image[ann["image_id"]].append(" ".join(self._words(ann["caption"])))
        return list(by_image), list(file_of.values()), list(by_image.values())

    def __len__(self) -> int:
        return len(self.image_ids)

    def __getitem__(self, idx: int) -> Instance:
        return Instance(filename=self.filenames[idx], captions=self.captions_with_image[idx],
                        **self.load_features(self.image_ids[idx]))
