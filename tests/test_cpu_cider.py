"""Native CIDEr-D (csrc/host_cider.cpp) against scores computed by the REAL reference
(tests/golden/cider.json, written by oracle/ref_harness/gen_golden_cider.py).  float64 arithmetic in the reference's
order of operations: the scores must agree to rounding -- TOL below; in this container they are bit-identical."""

import json

import numpy as np
import pytest

from openviic_b200.evaluation import Cider

from conftest import GOLDEN

TOL = 1e-12   # absolute, on scores of magnitude 0..10


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN / "cider.json", encoding="utf-8") as fh:
        return json.load(fh)


def _check(got, want):
    mean, scores = got
    assert scores.dtype == np.float64 and scores.shape == (len(want["scores"]),)
    assert np.abs(scores - np.array(want["scores"])).max() <= TOL
    assert abs(mean - want["mean"]) <= TOL


def test_reward_with_training_corpus(golden, cap_lib):
    cider = Cider(golden["corpus"])
    assert abs(cider.ref_len - np.log(len(golden["corpus"]))) == 0
    _check(cider.compute_score(golden["gts"], golden["gens"]), golden["with_corpus"])
    _check(cider.compute_score(golden["gts"], golden["gens"]), golden["with_corpus"])   # cached references, same result
    _check(cider.compute_score({"a": ["một người đang đi bộ"]}, {"a": ["một người đi bộ"]}), golden["single"])
    assert str(cider) == "CIDEr"


def test_batch_document_frequencies(golden, cap_lib):
    _check(Cider().compute_score(golden["gts"], golden["gens"]), golden["batch_only"])


def test_other_sigma(golden, cap_lib):
    """Only sigma can be varied against the reference: its cook_refs / cook_test ignore n (cider_scorer.py:26-45)
    and any n other than 4 raises IndexError there."""
    _check(Cider(golden["corpus"], sigma=3.0).compute_score(golden["gts"], golden["gens"]), golden["sigma3"])
    low = Cider(golden["corpus"], n=2).compute_score(golden["gts"], golden["gens"])[1]   # runs here; no reference value
    assert np.isfinite(low).all() and low.shape == (len(golden["gens"]),)


def test_thread_count_does_not_change_scores(golden, cap_lib):
    one = Cider(golden["corpus"], threads=1).compute_score(golden["gts"], golden["gens"])[1]
    many = Cider(golden["corpus"], threads=7).compute_score(golden["gts"], golden["gens"])[1]
    assert np.array_equal(one, many)


def test_reward_shape_for_the_self_critical_step(golden, cap_lib):
    """The trainer reshapes the scores to (images, beam) and subtracts the per-image mean
    (reference trainers/vi_trainer.py:143-146)."""
    scores = Cider(golden["corpus"]).compute_score(golden["gts"], golden["gens"])[1].astype(np.float32)
    reward = scores.reshape(-1, golden["beam"])
    assert reward.shape[0] * golden["beam"] == len(golden["gens"])
    assert np.isfinite(reward).all() and (reward >= 0).all()
    assert reward[0, 1] == reward[0].max()     # of image 0's beams, the one copied from a reference scores highest
    assert reward[0, 0] == 0 and reward[0, 3] == 0   # empty / all-unseen hypotheses score zero


def test_argument_errors(golden, cap_lib):
    cider = Cider(golden["corpus"])
    with pytest.raises(AssertionError):        # the reference asserts gts.keys() == res.keys()
        cider.compute_score({"a": ["x"]}, {"b": ["x"]})
    with pytest.raises(RuntimeError, match="no reference captions"):
        cider.compute_score({"a": []}, {"a": ["x"]})
    with pytest.raises(TypeError):
        cider.compute_score({"a": "x y"}, {"a": ["x"]})
    with pytest.raises(RuntimeError, match="n must be"):
        Cider(n=5)
