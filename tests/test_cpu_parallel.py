"""Multi-rank host logic on CPU: world_size-2 gloo, shard -> caption -> all-gather == single rank."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openviic_b200 import parallel
from oracle import caption_oracle as oracle
from helpers import golden, load_case


def test_shard_bounds_cover_the_batch_once():
    for total in (0, 1, 7, 256, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    r, w, _ = parallel.init_distributed("gloo")
    case, cfg, vocab, model, weights, field, feats, boxes = load_case("ort_trig")
    total = feats.shape[0] - 1                      # 3 images over 2 ranks: uneven shards
    feats, boxes = feats[:total], boxes[:total]
    lo, hi = parallel.shard_bounds(total, w, r)
    # the per-rank caption step is stood in for by the oracle (tests may run it; the product runs the engine)
    ids, lp = oracle.caption_beam_search(weights, cfg.MODEL, vocab, parallel.shard_tensor(feats, w, r),
                                         parallel.shard_tensor(boxes, w, r), beam=case["beam"], out_size=1)
    assert ids.shape[0] == hi - lo
    full_ids, full_lp = parallel.gather_captions(ids, lp, total)
    np.save(os.path.join(out_dir, f"ids_{r}.npy"), full_ids.numpy())
    np.save(os.path.join(out_dir, f"lp_{r}.npy"), full_lp.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_rank(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = golden("ort_trig")
    for r in range(2):
        ids = np.load(tmp_path / f"ids_{r}.npy")
        lp = np.load(tmp_path / f"lp_{r}.npy")
        assert ids.dtype == np.int64 and np.array_equal(ids, g["ids"][:3])     # images are independent units
        assert np.abs(lp - g["logp"][:3]).max() < 2e-5


def test_gather_is_identity_without_a_process_group():
    ids, lp = torch.arange(6).view(2, 3), torch.zeros(2, 3)
    out_ids, out_lp = parallel.gather_captions(ids, lp, 2)
    assert out_ids is ids and out_lp is lp


def _train_worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    r, w, _ = parallel.init_distributed("gloo")
    # a stand-in "model": loss = mean over valid tokens of the whole batch of (theta . x_token)^2
    g = torch.Generator().manual_seed(3)
    x = torch.randn(7, 5, generator=g)                 # 7 tokens over 2 ranks, uneven
    valid = torch.tensor([1, 1, 0, 1, 1, 1, 0], dtype=torch.bool)
    theta = torch.linspace(-1, 1, 5)
    lo, hi = parallel.shard_bounds(7, w, r)
    xs, vs = x[lo:hi], valid[lo:hi]
    inv = parallel.global_token_weight(vs.sum())
    pred = xs @ theta
    loss = ((pred ** 2) * vs).sum() * inv
    grad = ((2 * pred * vs).unsqueeze(1) * xs).sum(0) * inv
    parallel.sum_gradients_(grad, loss)
    np.save(os.path.join(out_dir, f"train_{r}.npy"), np.concatenate([[loss.item(), inv.item()], grad.numpy()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_equals_the_global_batch(tmp_path):
    """The training step's N > 1 path (parallel.global_token_weight + sum_gradients_), world size 2 on gloo, uneven
    shards: every rank ends up with the loss and gradient of the whole batch, not an average of per-rank means."""
    mp.spawn(_train_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(7, 5, generator=g)
    valid = torch.tensor([1, 1, 0, 1, 1, 1, 0], dtype=torch.bool)
    theta = torch.linspace(-1, 1, 5).requires_grad_(True)
    loss = (((x @ theta) ** 2) * valid).sum() / valid.sum()
    loss.backward()
    for r in range(2):
        got = np.load(tmp_path / f"train_{r}.npy")
        assert abs(got[0] - loss.item()) < 1e-6 and abs(got[1] - 1.0 / 5) < 1e-7
        assert np.abs(got[2:] - theta.grad.numpy()).max() < 1e-6
