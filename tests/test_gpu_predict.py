"""Host rows -> caption text through CaptionPredictor (native collate + engine + native decode) against the
reference's procedure for the same batch: the model's own ``beam_search`` on the zero-padded batch, then the
Python decode loop and duplicate collapse of the evaluation loop (reference trainers/vi_trainer.py:242-251)."""

import itertools

import pytest
import torch

from openviic_b200.data_utils import Vocab
from openviic_b200.predict import CaptionPredictor
from helpers import load_case, make_items

pytestmark = pytest.mark.gpu


def _reference_text(vocab, ids):
    out = []
    for vec in ids.tolist():
        words = []
        for idx in vec:
            if vocab.itos[idx] not in vocab.specials:
                words.append(vocab.itos[idx])
            if idx == vocab.eos_idx:
                break
        out.append(" ".join(k for k, _ in itertools.groupby(words)))
    return out


@pytest.mark.parametrize("name", ["std_region_A", "std_grid", "ort"])
def test_predictor_matches_model_and_reference_decode(name, device):
    case, cfg, vocab, model, weights, field, feats, boxes = load_case(name, device)
    b, n, beam = case["batch"], case["n"], case["beam"]
    feats = feats.to(torch.bfloat16).float()     # both sides then start from the same bf16 feature values
    valid = [int((feats[i].abs().sum(-1) != 0).sum()) for i in range(b)]
    assert all(not feats[i, v:].any() for i, v in enumerate(valid))   # padding is a suffix of all-zero rows
    rows = [feats[i, :v].numpy() for i, v in enumerate(valid)]
    box_rows = None if boxes is None else [boxes[i, :v].numpy() for i, v in enumerate(valid)]

    text_vocab = Vocab.from_itos(vocab.itos, vocab.max_caption_length)
    predictor = CaptionPredictor(model, text_vocab, max_batch=b, max_rows=n, beam_size=beam)
    got = predictor.predict(rows, box_rows)
    again = predictor.predict(rows, box_rows)    # second call replays the captured graph from the other staging slot

    padded_boxes = None
    if boxes is not None:                        # the reference's collate zero-pads boxes too
        padded_boxes = boxes.clone()
        for i, v in enumerate(valid):
            padded_boxes[i, v:] = 0
    ids, _ = model.beam_search(make_items(field, feats, padded_boxes, device), batch_size=b, beam_size=beam, out_size=1)
    torch.cuda.synchronize()
    want = _reference_text(text_vocab, ids.view(-1, vocab.max_caption_length).cpu())
    assert len(got) == b and all(isinstance(c, str) for c in got)
    assert got == want
    assert again == want


def test_predictor_smaller_last_batch(device):
    """The last batch of a dataset is smaller than the reservation: same captions as the full batch gave."""
    case, cfg, vocab, model, weights, field, feats, boxes = load_case("std_grid", device)
    b, n, beam = case["batch"], case["n"], case["beam"]
    rows = [feats[i].numpy() for i in range(b)]
    predictor = CaptionPredictor(model, Vocab.from_itos(vocab.itos, vocab.max_caption_length), b, n, beam)
    full = predictor.predict(rows)
    assert predictor.predict(rows[: b // 2]) == full[: b // 2]
    assert predictor.predict(rows) == full
