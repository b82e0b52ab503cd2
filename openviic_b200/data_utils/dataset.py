"""Annotation + pre-computed feature datasets feeding the caption path
(reference: data_utils/dataset.py:12-72 ``FeatureDataset``, :74-132 ``DictionaryDataset``).

On disk, as the reference expects it: one COCO-style JSON (``images`` with ``id`` / ``file_name``, ``annotations``
with ``image_id`` / ``caption``) and one ``{image_id}.npy`` per image under ``config.FEATURE_PATH.FEATURES`` holding a
pickled dict of arrays (``region_features`` (n, D), ``region_boxes`` (n, 4), ``grid_features`` ...), read with
``np.load(..., allow_pickle=True)[()]``.  Samples are ``Instance`` objects; batches are built by
``data_utils.collate_fn`` (the reference's) or, on the fast path, by ``FeatureBatcher`` from the per-image arrays.
"""

from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Tuple

import numpy as np
import torch
from torch.utils import data

from ..utils.instance import Instance
from .utils import preprocess_caption


def _load_features(root: str, image_id) -> Dict[str, Any]:
    return np.load(os.path.join(root, f"{image_id}.npy"), allow_pickle=True)[()]


class FeatureDataset(data.Dataset):
    """One sample per ANNOTATION: teacher-forcing tokens + the image's features (XE training / validation loss)."""

    def __init__(self, json_path: str, vocab, config) -> None:
        super().__init__()
        with open(json_path, "r", encoding="utf-8") as fh:
            json_data = json.load(fh)
        self.vocab = vocab
        self.annotations = self.load_json(json_data)
        self.image_features_path = config.FEATURE_PATH.FEATURES

    def load_json(self, json_data: Dict) -> List[Dict]:
        # the reference scans the image list per annotation (data_utils/dataset.py:29-42); an id -> file name table
        # gives the same records (an annotation whose image is missing raises KeyError here; the reference silently
        # repeats the previous annotation)
        filenames = {}
        for image in json_data["images"]:
            filenames.setdefault(image["id"], image["file_name"])
        return [{"caption": preprocess_caption(ann["caption"], self.vocab.tokenizer),
                 "image_id": ann["image_id"],
                 "filename": filenames[ann["image_id"]]} for ann in json_data["annotations"]]

    def load_features(self, image_id) -> Dict[str, Any]:
        return _load_features(self.image_features_path, image_id)

    @property
    def captions(self) -> List[List[str]]:
        return [ann["caption"] for ann in self.annotations]

    def __getitem__(self, idx: int) -> Instance:
        item = self.annotations[idx]
        caption = self.vocab.encode_caption(item["caption"])
        shifted = torch.full_like(caption, self.vocab.padding_idx)
        shifted[:-1] = caption[1:]
        caption = torch.where(caption == self.vocab.eos_idx, self.vocab.padding_idx, caption)   # the input never holds eos
        return Instance(caption_tokens=caption, shifted_right_caption_tokens=shifted, **self.load_features(item["image_id"]))

    def __len__(self) -> int:
        return len(self.annotations)


class DictionaryDataset(data.Dataset):
    """One sample per IMAGE with all of its reference captions (evaluation, self-critical training)."""

    def __init__(self, json_path: str, vocab, config) -> None:
        super().__init__()
        with open(json_path, "r", encoding="utf-8") as fh:
            json_data = json.load(fh)
        self.vocab = vocab
        self.image_ids, self.filenames, self.captions_with_image = self.load_json(json_data)
        self.image_features_path = config.FEATURE_PATH.FEATURES

    def load_json(self, json_data: Dict) -> Tuple[List, List[str], List[List[str]]]:
        examples: Dict[Any, List[str]] = {}
        filenames: Dict[Any, str] = {}
        for image in json_data["images"]:
            examples[image["id"]] = []
            filenames[image["id"]] = image["file_name"]
        for ann in json_data["annotations"]:
            examples[ann["image_id"]].append(" ".join(preprocess_caption(ann["caption"], self.vocab.tokenizer)))
        return list(examples.keys()), list(filenames.values()), list(examples.values())

    def load_features(self, image_id) -> Dict[str, Any]:
        return _load_features(self.image_features_path, image_id)

    def __len__(self) -> int:
        return len(self.image_ids)

    def __getitem__(self, idx: int) -> Instance:
        return Instance(filename=self.filenames[idx], captions=self.captions_with_image[idx],
                        **self.load_features(self.image_ids[idx]))
