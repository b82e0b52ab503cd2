"""Host-side logic of the training step (T1) that needs no GPU: the schedule, the synthetic captions' format, the
counter-based dropout mask's restatement, and the trainer's refusal to run anywhere but on CUDA."""

import math

import numpy as np
import pytest
import torch

import openviic_b200 as ov
from openviic_b200 import synthetic
from openviic_b200.training import XETrainer, noam_factor
from oracle import caption_oracle as oracle


def test_noam_schedule_is_the_references_lambda():
    """trainers/base_trainer.py:114-117: (d_model ** -.5) * min((step + 1) ** -.5, (step + 1) * warmup ** -1.5)."""
    for d_model, warmup in ((512, 10000), (512, 100), (256, 7)):
        for step in (0, 1, 5, warmup - 2, warmup - 1, warmup, 3 * warmup):
            s = step + 1
            want = (d_model ** -.5) * min(s ** -.5, s * warmup ** -1.5)
            assert noam_factor(step, d_model, warmup) == want == oracle.noam_factor(step, d_model, warmup)
    assert noam_factor(99, 512, 100) == max(noam_factor(s, 512, 100) for s in range(400))   # the peak is at step = warmup


def test_synthetic_captions_have_the_datasets_format():
    """data_utils/dataset.py:56-61: input row = <bos> words <pad>.. (its <eos> replaced by <pad>), targets = the encoded
    caption shifted left by one (words <eos> <pad>..)."""
    tokens, targets = synthetic.synth_captions(32, 20, 1000, seed=5)
    assert tokens.dtype == targets.dtype == torch.int64 and tokens.shape == targets.shape == (32, 20)
    assert (tokens[:, 0] == 1).all() and not (tokens == 2).any()
    for row_in, row_out in zip(tokens.tolist(), targets.tolist()):
        n_words = sum(1 for t in row_in if t >= 4)
        assert 3 <= n_words <= 18
        assert row_in[1:1 + n_words] == row_out[:n_words] and row_out[n_words] == 2
        assert all(t == 0 for t in row_in[1 + n_words:]) and all(t == 0 for t in row_out[n_words + 1:])


def test_dropout_mask_restatement_is_a_deterministic_fair_coin():
    n = 200000
    a = oracle.dropout_keep(n, 1234, "encoder.layers.0.pwff.dropout", 0.1)
    assert torch.equal(a, oracle.dropout_keep(n, 1234, "encoder.layers.0.pwff.dropout", 0.1))
    for other in (oracle.dropout_keep(n, 1235, "encoder.layers.0.pwff.dropout", 0.1),
                  oracle.dropout_keep(n, 1234, "encoder.layers.1.pwff.dropout", 0.1)):
        agree = (a == other).float().mean().item()          # independent masks agree on 0.9^2 + 0.1^2 = 0.82 of the elements
        assert abs(agree - 0.82) < 0.01
    for p in (0.1, 0.5):
        keep = oracle.dropout_keep(n, 7, "vision_embedding.dropout", p).float()
        assert abs(keep.mean().item() - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n)
        # no short-range structure: neighbouring elements are kept independently
        assert abs((keep[1:] * keep[:-1]).mean().item() - (1 - p) ** 2) < 0.01
    # the hash in plain Python integers (what csrc/train.cu dropout_hash computes in uint32 registers)
    def hash32(i, seed, site):
        m = 0xFFFFFFFF
        x = (i * 0x9E3779B1 + seed * 0x85EBCA77 + site * 0xC2B2AE3D) & m
        x ^= x >> 15
        x = (x * 0x2C1B3C6D) & m
        x ^= x >> 12
        x = (x * 0x297A2D39) & m
        x ^= x >> 15
        return x
    site = oracle.dropout_site("vision_embedding.dropout")
    want = [hash32(i, 0xFFFFFFF0, site) >= oracle.dropout_threshold(0.3) for i in range(64)]
    assert oracle.dropout_keep(64, 0xFFFFFFF0, "vision_embedding.dropout", 0.3).tolist() == want


def test_trainer_refuses_the_cpu():
    cfg = ov.get_config("standard_transformer.yaml")
    cfg.MODEL.DEVICE = "cpu"
    model = ov.build_model(cfg.MODEL, synthetic.SyntheticVocab(50, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        XETrainer(model)
