"""Host glue either side of the path against outputs of the REAL reference
(tests/golden/text_glue.*, written by oracle/ref_harness/gen_golden_text.py): caption pre-processing,
vocabulary construction / encode / decode, the evaluation loop's duplicate collapse, and the feature collate.
Everything is host code: the natively decoded strings and collated batches must be identical, not close."""

import itertools
import json

import numpy as np
import pytest
import torch

from openviic_b200.configs.utils import CfgNode
from openviic_b200.data_utils import FeatureBatcher, Vocab, collate_fn, preprocess_caption
from openviic_b200.utils.instance import Instance

from conftest import GOLDEN


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN / "text_glue.json", encoding="utf-8") as fh:
        return json.load(fh), np.load(GOLDEN / "text_glue.npz")


def _vocab_config(tmp_path, annotations, min_freq):
    paths = {}
    for split, caps in annotations.items():
        paths[split] = str(tmp_path / f"{split.lower()}.json")
        with open(paths[split], "w", encoding="utf-8") as fh:
            json.dump({"annotations": [{"image_id": i, "caption": c} for i, c in enumerate(caps)]}, fh, ensure_ascii=False)
    return CfgNode({"MIN_FREQ": min_freq,
                    "VOCAB": {"TOKENIZER": None, "WORD_EMBEDDING": None, "WORD_EMBEDDING_CACHE": None,
                              "BOS_TOKEN": "<bos>", "EOS_TOKEN": "<eos>", "PAD_TOKEN": "<pad>", "UNK_TOKEN": "<unk>",
                              "USE_MAPPING": False, "PRETRAINED_LANGUAGE_MODEL": None},
                    "JSON_PATH": paths})


def test_preprocess_caption_matches_reference(golden):
    g, _ = golden
    for probe, want in zip(g["probes"], g["vocabs"][0]["preprocessed"]):
        assert preprocess_caption(probe, None) == want
    assert preprocess_caption("A,b", tokenizer=lambda s: s.replace("a", "x y")) == ["x", "y", ",", "b"]


@pytest.mark.parametrize("which", [0, 1])
def test_vocab_construction_and_encode(golden, which, tmp_path):
    g, arrays = golden
    want = g["vocabs"][which]
    vocab = Vocab(_vocab_config(tmp_path, g["annotations"], want["min_freq"]))
    assert vocab.itos == want["itos"]                      # specials, then falling frequency, ties alphabetical
    assert dict(vocab.freqs) == want["freqs"]
    assert vocab.max_caption_length == want["max_caption_length"]
    assert [vocab.padding_idx, vocab.bos_idx, vocab.eos_idx, vocab.unk_idx] == want["specials"]
    assert len(vocab) == len(want["itos"]) and vocab.stoi["<unk>"] == vocab.unk_idx
    encoded = torch.stack([vocab.encode_caption(words) for words in want["encoded_inputs"]])
    assert encoded.dtype == torch.int64
    assert np.array_equal(encoded.numpy(), arrays[f"encoded_minfreq{want['min_freq']}"])
    with pytest.raises(IndexError):                        # the reference overruns its vector the same way
        vocab.encode_caption(["một"] * vocab.max_caption_length)


@pytest.mark.parametrize("which", [0, 1])
def test_decode_caption_matches_reference(golden, which, cap_lib):
    g, arrays = golden
    want = g["vocabs"][which]
    vocab = Vocab.from_itos(want["itos"], want["max_caption_length"])
    ids = torch.from_numpy(arrays[f"ids_minfreq{want['min_freq']}"])
    assert vocab.decode_caption(ids) == want["decoded_joined"]
    assert vocab.decode_caption(ids, join_words=False) == want["decoded_words"]
    assert vocab.decode_predictions(ids) == want["decoded_collapsed"]
    assert vocab.decode_caption(ids, join_words=True, collapse_repeats=True) == want["decoded_collapsed"]
    # the shape the evaluation loop passes: (B, out_size, T) flattened, possibly a non-contiguous view, int32 ids
    assert vocab.decode_caption(ids.view(8, 8, -1).transpose(0, 1).reshape(-1, ids.shape[1])) == \
        [want["decoded_joined"][(i % 8) * 8 + i // 8] for i in range(64)]
    assert vocab.decode_caption(ids.to(torch.int32).numpy()) == want["decoded_joined"]
    assert vocab.decode_caption(ids[:0]) == []


def test_decode_caption_edge_cases(cap_lib):
    itos = ["<pad>", "<bos>", "<eos>", "<unk>", "a", "bb", "ccc", "đường"]
    vocab = Vocab.from_itos(itos, 6)
    ids = torch.tensor([[1, 4, 4, 5, 2, 6],      # repeat, eos stops the caption
                        [1, 3, 0, 1, 7, 7],      # specials skipped, no eos
                        [2, 4, 4, 4, 4, 4],      # eos first: empty
                        [4, 5, 4, 5, 5, 4]])     # alternating words are not repeats
    assert vocab.decode_caption(ids) == ["a a bb", "đường đường", "", "a bb a bb bb a"]
    assert vocab.decode_predictions(ids) == ["a bb", "đường", "", "a bb a bb a"]
    assert vocab.decode_caption(ids, join_words=False)[2] == []
    with pytest.raises(IndexError):              # the reference: itos[idx] -> IndexError
        vocab.decode_caption(torch.tensor([[1, 99, 2, 0, 0, 0]]))
    with pytest.raises(ValueError):
        vocab.decode_caption(torch.tensor([1, 2, 3]))
    # a vocabulary whose words str.split() would cut: the collapse follows the reference (over the split words)
    odd = Vocab.from_itos(["<pad>", "<bos>", "<eos>", "<unk>", "x y", "y", "x"], 6)
    row = torch.tensor([[1, 4, 5, 6, 6, 2]])     # "x y" "y" "x" "x"
    words = odd.decode_caption(row, join_words=False)[0]
    assert words == ["x", "y", "y", "x", "x"]
    assert odd.decode_predictions(row) == [" ".join(k for k, _ in itertools.groupby(words))] == ["x y x"]


def test_decode_at_scale_against_the_loop(cap_lib):
    """8192 captions of the bench's shape (V 10201, T 20) against a restatement of the reference loop."""
    vocab = Vocab.from_itos(["<pad>", "<bos>", "<eos>", "<unk>"] + [f"w{i}" for i in range(4, 10201)], 20)
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 40, size=(8192, 20))   # a small id range: many eos, specials and repeats
    specials = set(vocab.specials)
    want = []
    for row in ids.tolist():
        words = []
        for idx in row:
            if vocab.itos[idx] not in specials:
                words.append(vocab.itos[idx])
            if idx == vocab.eos_idx:
                break
        want.append(" ".join(k for k, _ in itertools.groupby(words)))
    assert vocab.decode_predictions(torch.from_numpy(ids)) == want


def test_instance_list_collate_matches_reference(golden):
    g, arrays = golden
    lengths = arrays["collate_lengths"].tolist()
    bounds = np.cumsum([0] + lengths)
    feats = [arrays["collate_feats_in"][a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    boxes = [arrays["collate_boxes_in"][a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    batch = collate_fn([Instance(region_features=f, region_boxes=b, filename=f"img{i}.jpg")
                        for i, (f, b) in enumerate(zip(feats, boxes))])
    assert batch.batch_size == g["collate_batch_size"]
    assert list(batch.filename) == g["collate_filenames"]
    assert batch.region_features.dtype == torch.float32
    assert np.array_equal(batch.region_features.numpy(), arrays["collate_feats_out"])
    assert np.array_equal(batch.region_boxes.numpy(), arrays["collate_boxes_out"])


def test_feature_batcher_matches_reference_collate(golden, cap_lib):
    """The native collate = the reference's padded fp32 batch, rounded to bf16 the way torch rounds."""
    g, arrays = golden
    lengths = arrays["collate_lengths"].tolist()
    bounds = np.cumsum([0] + lengths)
    feats = [arrays["collate_feats_in"][a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    boxes = [arrays["collate_boxes_in"][a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    batcher = FeatureBatcher(max_batch=8, max_rows=12, width=16, box_width=4, slots=2, threads=3)
    out, out_boxes = batcher.collate(feats, boxes)
    want = torch.from_numpy(arrays["collate_feats_out"])
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == tuple(want.shape)
    assert torch.equal(out, want.to(torch.bfloat16))
    assert torch.equal(out_boxes, torch.from_numpy(arrays["collate_boxes_out"]))
    # fixed shape (one CUDA graph for every batch): longer padding, same leading rows; the second slot is used
    out12, _ = batcher.collate([torch.from_numpy(f) for f in feats], pad_to=12)
    assert out12.data_ptr() != out.data_ptr() and tuple(out12.shape) == (5, 12, 16)
    assert torch.equal(out12[:, : want.shape[1]], want.to(torch.bfloat16)) and not out12[:, want.shape[1]:].any()
    with pytest.raises(ValueError):
        batcher.collate(feats, pad_to=13)
    with pytest.raises(ValueError):
        batcher.collate(feats, boxes[:-1] + [boxes[-1][:-1]])
    with pytest.raises(ValueError):
        batcher.collate([f[:, :8] for f in feats])


@pytest.mark.parametrize("width", [16, 24, 5])   # whole vectors; unaligned rows with a scalar tail; scalar only
def test_bf16_rounding_is_torchs(width, cap_lib):
    """Round-to-nearest-even on every class of value: ties, subnormals, overflow to inf, signed zeros, inf, NaN."""
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 32, size=1 << 16, dtype=np.uint64).astype(np.uint32)
    special = np.array([0x00000000, 0x80000000, 0x3F808000, 0x3F818000, 0x3F80FFFF, 0x7F7FFFFF, 0xFF7FFFFF,
                        0x7F800000, 0xFF800000, 0x7FC00000, 0xFFC00001, 0x7F800001, 0x00000001, 0x00008000,
                        0x00018000, 0x807FFFFF], dtype=np.uint32)
    x = np.concatenate([special, bits]).view(np.float32)
    x = x[: x.size // width * width].reshape(-1, width)
    out, _ = FeatureBatcher(max_batch=1, max_rows=x.shape[0], width=width, threads=4).collate([x])
    want = torch.from_numpy(x).to(torch.bfloat16)
    got_bits, want_bits = out[0].view(torch.int16), want.view(torch.int16)
    nan = torch.isnan(want)
    assert torch.equal(got_bits[~nan], want_bits[~nan])
    assert torch.isnan(out[0][nan]).all()


def _dataset_fixture(golden, tmp_path):
    """The annotation JSON and the per-image .npy feature dicts the golden generator fed to the reference."""
    g, arrays = golden
    d = g["dataset"]
    feat_dir = tmp_path / "features"
    feat_dir.mkdir()
    bounds = np.cumsum([0] + d["rows_per_image"])
    for k, image in enumerate(d["images"]):
        np.save(feat_dir / f"{image['id']}.npy",
                {"region_features": arrays["dataset_feats_in"][bounds[k]:bounds[k + 1]],
                 "region_boxes": arrays["dataset_boxes_in"][bounds[k]:bounds[k + 1]]}, allow_pickle=True)
    json_path = tmp_path / "dataset.json"
    with open(json_path, "w", encoding="utf-8") as fh:
        json.dump({"images": d["images"], "annotations": d["annotations"]}, fh, ensure_ascii=False)
    vocab = Vocab(_vocab_config(tmp_path, g["annotations"], 2))     # the generator used the MIN_FREQ 2 vocabulary
    return d, arrays, str(json_path), vocab, CfgNode({"FEATURE_PATH": {"FEATURES": str(feat_dir)}})


def test_feature_dataset_matches_reference(golden, tmp_path):
    from openviic_b200.data_utils import FeatureDataset
    d, arrays, json_path, vocab, cfg = _dataset_fixture(golden, tmp_path)
    ds = FeatureDataset(json_path, vocab, cfg)
    assert len(ds) == d["feature_len"]
    assert ds.captions == d["feature_captions"]
    samples = [ds[i] for i in range(len(ds))]
    assert [list(s.keys()) for s in samples] == d["feature_fields"]
    assert np.array_equal(torch.stack([s.caption_tokens for s in samples]).numpy(), arrays["dataset_caption_tokens"])
    assert np.array_equal(torch.stack([s.shifted_right_caption_tokens for s in samples]).numpy(),
                          arrays["dataset_shifted_tokens"])
    assert [s.region_features.shape[0] for s in samples] == arrays["dataset_sample_rows"].tolist()
    assert not (torch.stack([s.caption_tokens for s in samples]) == vocab.eos_idx).any()


def test_dictionary_dataset_and_its_batches_match_reference(golden, tmp_path, cap_lib):
    from openviic_b200.data_utils import DictionaryDataset
    d, arrays, json_path, vocab, cfg = _dataset_fixture(golden, tmp_path)
    ds = DictionaryDataset(json_path, vocab, cfg)
    assert len(ds) == d["dictionary_len"]
    assert list(ds.image_ids) == d["dictionary_image_ids"]
    assert list(ds.filenames) == d["dictionary_filenames"]
    assert ds.captions_with_image == d["dictionary_captions"]
    samples = [ds[i] for i in range(len(ds))]
    batch = collate_fn(samples)                                      # the reference's batch
    assert list(batch.keys()) == d["batch_fields"] and batch.batch_size == d["batch_size"]
    assert list(batch.filename) == d["batch_filename"] and list(batch.captions) == d["batch_captions"]
    assert np.array_equal(batch.region_features.numpy(), arrays["dataset_batch_feats"])
    assert np.array_equal(batch.region_boxes.numpy(), arrays["dataset_batch_boxes"])
    # the fast path: the same samples through the native collate (bf16 features, fp32 boxes, pinned when a GPU is present)
    batcher = FeatureBatcher(max_batch=8, max_rows=16, width=16, box_width=4)
    feats, boxes = batcher.collate([s.region_features for s in samples], [s.region_boxes for s in samples])
    assert torch.equal(feats, torch.from_numpy(arrays["dataset_batch_feats"]).to(torch.bfloat16))
    assert torch.equal(boxes, torch.from_numpy(arrays["dataset_batch_boxes"]))


def test_get_predictions_loop_with_a_scripted_predictor(golden, tmp_path, cap_lib):
    """The evaluation loop's bookkeeping (keys, grouping, overlap protocol, CIDEr over the whole set) with a scripted
    stand-in for the GPU predictor; the GPU leg is tests/test_gpu_predict.py."""
    import itertools as it_
    from openviic_b200.data_utils import DictionaryDataset
    from openviic_b200.evaluation import Cider
    from openviic_b200.predict import evaluate_metrics, get_predictions
    d, arrays, json_path, vocab, cfg = _dataset_fixture(golden, tmp_path)
    ds = DictionaryDataset(json_path, vocab, cfg)

    class Scripted:
        def __init__(self):
            self.inflight, self.calls = None, []

        def submit(self, features, boxes):
            assert self.inflight is None and boxes is not None and len(boxes) == len(features)
            assert all(f.shape[0] == b.shape[0] for f, b in zip(features, boxes))
            self.calls.append(len(features))
            self.inflight = ["caption with %d rows" % f.shape[0] for f in features]

        def collect(self):
            out, self.inflight = self.inflight, None
            return out

    pred = Scripted()
    out = get_predictions(pred, ds, batch_size=3, get_scores=False)   # image 8 has no reference captions
    assert "CIDEr" not in out
    assert pred.calls == [3, 1] and pred.inflight is None
    assert [r["image_id"] for r in out["results"]] == [d["dictionary_image_ids"][:3], d["dictionary_image_ids"][3:]]
    assert out["results"][0]["filename"] == d["dictionary_filenames"][:3]
    assert list(out["results"][0]["gens"]) == ["0_0", "0_1", "0_2"] and list(out["results"][1]["gts"]) == ["1_0"]
    assert out["results"][0]["gens"]["0_1"] == "caption with %d rows" % d["rows_per_image"][1]
    assert out["results"][0]["gts"]["0_0"] == d["dictionary_captions"][0]
    # an image without reference captions cannot be scored (the reference's scorer divides by zero there): drop it
    scored = [i for i, caps in enumerate(d["dictionary_captions"]) if caps]

    class Subset:
        image_ids = [ds.image_ids[i] for i in scored]

        def __len__(self):
            return len(scored)

        def __getitem__(self, i):
            return ds[scored[i]]

    scores = evaluate_metrics(Scripted(), Subset(), batch_size=2)
    gts = {"%d_%d" % (k // 2, k % 2): d["dictionary_captions"][i] for k, i in enumerate(scored)}
    gens = {"%d_%d" % (k // 2, k % 2): ["caption with %d rows" % d["rows_per_image"][i]] for k, i in enumerate(scored)}
    assert scores == {"CIDEr": float(Cider().compute_score(gts, gens)[0])}
