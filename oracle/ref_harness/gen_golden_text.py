"""Generate tests/golden/text_glue.json + text_glue.npz by running the REAL reference's host glue on CPU.

Test infrastructure (build container only; the reference cannot travel to the GPU box).  Pins SURVEY.md
section 8f row 2 -- the code either side of the caption path:

  * ``data_utils.utils.preprocess_caption``  (reference data_utils/utils.py:57-80)
  * ``data_utils.vocab.Vocab``: construction from annotation files, ``encode_caption``, ``decode_caption``
    (reference data_utils/vocab.py:16-122)
  * the evaluation loop's duplicate collapse, ``' '.join(k for k, g in itertools.groupby(words))``
    (reference trainers/vi_trainer.py:251) -- re-stated here in one line because it is inlined in the trainer
  * ``InstanceList`` collate of ragged per-image features (reference utils/instance.py:32-55,156-171)
  * ``FeatureDataset`` / ``DictionaryDataset`` over an annotation JSON + per-image ``.npy`` feature dicts
    (reference data_utils/dataset.py:12-132), and a collated ``DictionaryDataset`` batch

Everything recorded is an OUTPUT OF THE REFERENCE; the inputs (annotation files, id matrices, feature rows) are
stored next to it so that the tests rebuild them without the reference.

usage:  python oracle/ref_harness/gen_golden_text.py
"""

from __future__ import annotations

import itertools
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path(os.environ.get("OPENVIIC_REFERENCE", "/root/reference"))
sys.path.insert(0, str(HERE / "shims"))
sys.path.insert(0, str(REFERENCE))

import builders  # noqa: E402,F401  (must come first: data_utils.vocab <-> trainers import each other)
from data_utils.dataset import DictionaryDataset as RefDictionaryDataset, FeatureDataset as RefFeatureDataset  # noqa: E402
from data_utils.utils import collate_fn as ref_collate_fn, preprocess_caption as ref_preprocess  # noqa: E402
from data_utils.vocab import Vocab as RefVocab  # noqa: E402
from utils.instance import Instance as RefInstance  # noqa: E402
from yacs.config import CfgNode  # noqa: E402  (the shim)

# Captions in the style of UIT-OpenViIC (Vietnamese, lower/upper case, punctuation of the reference's list,
# curly quotes, repeated words, frequency ties, a word that only appears once -> dropped by MIN_FREQ 2).
TRAIN = [
    "Một người đàn ông đang đi bộ trên đường phố.",
    "Một người phụ nữ đang bán hàng ở chợ, xung quanh có nhiều người!",
    "Hai người đang ngồi (trên ghế) trước cửa hàng: \"Tạp hoá Số 5\".",
    "Một chiếc xe máy màu đỏ đang đậu trước cửa hàng; giá 5$/ngày?",
    "Những người đang đi đi lại lại trên đường phố.",
    "Một người đàn ông đang bán hàng ở chợ.",
    "Biển hiệu “Cà phê & Trà” treo trước cửa hàng [mới mở]...",
    "Một người phụ nữ đang đi xe máy trên đường phố * đông đúc *",
]
DEV = [
    "Một người đàn ông đang ngồi trước cửa hàng.",
    "Hai chiếc xe máy đang đậu trên đường phố, một chiếc màu đỏ.",
]
TEST = [
    "Một người phụ nữ đang ngồi trên ghế ở chợ.",
    "Xe máy, xe máy và xe máy: đường phố đông đúc!",
]
PROBES = [   # preprocess_caption / encode_caption probes, some with words outside the vocabulary
    "Một người đàn ông đang lái xe buýt trên đường phố.",
    "\"Tạp hoá\" (số 5) đang mở cửa; có nhiều người... đang mua hàng?",
    "Hai người phụ nữ đang đi bộ, một người đang bán hàng!",
    "   Khoảng   trắng   thừa   ",
    "",
    "Giá: 5$/ngày & 10$/tuần * [khuyến mãi] 'hôm nay'",
]


def main() -> None:
    tmp = Path(tempfile.mkdtemp(prefix="openviic_text_"))
    paths = {}
    for split, caps in (("TRAIN", TRAIN), ("DEV", DEV), ("TEST", TEST)):
        paths[split] = str(tmp / f"{split.lower()}.json")
        with open(paths[split], "w", encoding="utf-8") as fh:
            json.dump({"annotations": [{"image_id": i, "caption": c} for i, c in enumerate(caps)]}, fh, ensure_ascii=False)

    golden = {"annotations": {"TRAIN": TRAIN, "DEV": DEV, "TEST": TEST}, "probes": PROBES, "vocabs": []}
    arrays = {}
    for min_freq in (1, 2):
        cfg = CfgNode({"MIN_FREQ": min_freq,
                       "VOCAB": {"TOKENIZER": None, "WORD_EMBEDDING": None, "WORD_EMBEDDING_CACHE": None,
                                 "BOS_TOKEN": "<bos>", "EOS_TOKEN": "<eos>", "PAD_TOKEN": "<pad>", "UNK_TOKEN": "<unk>",
                                 "USE_MAPPING": False, "PRETRAINED_LANGUAGE_MODEL": None},
                       "JSON_PATH": paths})
        vocab = RefVocab(cfg)
        V, T = len(vocab), vocab.max_caption_length
        rng = np.random.default_rng(100 + min_freq)
        ids = rng.integers(0, V, size=(64, T), dtype=np.int64)
        ids[:, 0] = vocab.bos_idx                                  # what beam search emits first
        for r in range(0, 64, 2):                                  # an eos somewhere in every other row
            ids[r, rng.integers(1, T)] = vocab.eos_idx
        for r in range(0, 64, 3):                                  # runs of equal words
            at = rng.integers(1, T - 3)
            ids[r, at:at + 3] = ids[r, at]
        ids[5, 3:6] = [vocab.unk_idx, vocab.padding_idx, vocab.bos_idx]   # specials in the middle are skipped
        ids[6, :] = vocab.eos_idx                                  # empty caption
        ids[7, 1:] = ids[7, 1]                                     # one word repeated to the end, no eos
        words = vocab.decode_caption(torch.from_numpy(ids), join_words=False)
        probes_tok = [ref_preprocess(p, None) for p in PROBES]
        fitting = [w for w in probes_tok if len(w) + 2 <= T]
        golden["vocabs"].append({
            "min_freq": min_freq,
            "itos": list(vocab.itos),
            "freqs": dict(vocab.freqs),
            "max_caption_length": T,
            "specials": [vocab.padding_idx, vocab.bos_idx, vocab.eos_idx, vocab.unk_idx],
            "preprocessed": probes_tok,
            "encoded_inputs": fitting,
            "decoded_joined": vocab.decode_caption(torch.from_numpy(ids), join_words=True),
            "decoded_words": words,
            # reference trainers/vi_trainer.py:251
            "decoded_collapsed": [" ".join(k for k, _ in itertools.groupby(w)) for w in words],
        })
        arrays[f"ids_minfreq{min_freq}"] = ids
        arrays[f"encoded_minfreq{min_freq}"] = torch.stack([vocab.encode_caption(w) for w in fitting]).numpy()

    # InstanceList collate: ragged region features + boxes (fp32), one image already at the longest length
    rng = np.random.default_rng(7)
    lengths = [5, 9, 1, 9, 7]
    feats = [rng.standard_normal((n, 16)).astype(np.float32) for n in lengths]
    boxes = [rng.random((n, 4)).astype(np.float32) for n in lengths]
    batch = ref_collate_fn([RefInstance(region_features=f, region_boxes=b, filename=f"img{i}.jpg")
                            for i, (f, b) in enumerate(zip(feats, boxes))])
    arrays["collate_lengths"] = np.array(lengths)
    arrays["collate_feats_in"] = np.concatenate(feats, 0)
    arrays["collate_boxes_in"] = np.concatenate(boxes, 0)
    arrays["collate_feats_out"] = batch.region_features.numpy()
    arrays["collate_boxes_out"] = batch.region_boxes.numpy()
    golden["collate_filenames"] = list(batch.filename)
    golden["collate_batch_size"] = int(batch.batch_size)

    # datasets: 4 images (ids not in order, one without captions last), 7 annotations, per-image .npy feature dicts
    feat_dir = tmp / "features"
    feat_dir.mkdir()
    image_ids = [17, 3, 42, 8]
    rows_per_image = [6, 9, 4, 9]
    images = [{"id": i, "file_name": f"{i:06d}.jpg"} for i in image_ids]
    ann_images = [17, 3, 17, 42, 3, 17, 42]
    annotations = [{"image_id": i, "caption": TRAIN[k]} for k, i in enumerate(ann_images)]
    ds_json = str(tmp / "dataset.json")
    with open(ds_json, "w", encoding="utf-8") as fh:
        json.dump({"images": images, "annotations": annotations}, fh, ensure_ascii=False)
    ds_feats, ds_boxes = {}, {}
    for i, n in zip(image_ids, rows_per_image):
        ds_feats[i] = rng.standard_normal((n, 16)).astype(np.float32)
        ds_boxes[i] = rng.random((n, 4)).astype(np.float32)
        np.save(feat_dir / f"{i}.npy", {"region_features": ds_feats[i], "region_boxes": ds_boxes[i]}, allow_pickle=True)
    ds_cfg = CfgNode({"FEATURE_PATH": {"FEATURES": str(feat_dir)}})
    fds = RefFeatureDataset(ds_json, vocab, ds_cfg)          # `vocab`: the MIN_FREQ 2 vocabulary built above
    samples = [fds[i] for i in range(len(fds))]
    dds = RefDictionaryDataset(ds_json, vocab, ds_cfg)
    dict_batch = ref_collate_fn([dds[i] for i in range(len(dds))])
    golden["dataset"] = {
        "images": images, "annotations": annotations, "rows_per_image": rows_per_image,
        "feature_len": len(fds), "feature_captions": fds.captions, "feature_fields": [list(s.keys()) for s in samples],
        "dictionary_len": len(dds), "dictionary_image_ids": list(dds.image_ids), "dictionary_filenames": list(dds.filenames),
        "dictionary_captions": dds.captions_with_image,
        "batch_fields": list(dict_batch.keys()), "batch_filename": list(dict_batch.filename),
        "batch_captions": list(dict_batch.captions), "batch_size": int(dict_batch.batch_size),
    }
    arrays["dataset_feats_in"] = np.concatenate([ds_feats[i] for i in image_ids], 0)
    arrays["dataset_boxes_in"] = np.concatenate([ds_boxes[i] for i in image_ids], 0)
    arrays["dataset_caption_tokens"] = torch.stack([s.caption_tokens for s in samples]).numpy()
    arrays["dataset_shifted_tokens"] = torch.stack([s.shifted_right_caption_tokens for s in samples]).numpy()
    arrays["dataset_sample_rows"] = np.array([s.region_features.shape[0] for s in samples])
    arrays["dataset_batch_feats"] = dict_batch.region_features.numpy()
    arrays["dataset_batch_boxes"] = dict_batch.region_boxes.numpy()

    out_dir = REPO / "tests" / "golden"
    with open(out_dir / "text_glue.json", "w", encoding="utf-8") as fh:
        json.dump(golden, fh, ensure_ascii=False, indent=1)
    np.savez_compressed(out_dir / "text_glue.npz", **arrays)
    print(f"wrote {out_dir / 'text_glue.json'} and text_glue.npz "
          f"(vocab sizes {[len(v['itos']) for v in golden['vocabs']]}, T {golden['vocabs'][0]['max_caption_length']})")


if __name__ == "__main__":
    main()
