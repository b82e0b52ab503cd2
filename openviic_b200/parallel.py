"""Data-parallel sharding of the image batch: one process per GPU, no data-path collective.

Images are independent units (nothing in encoder, decoder or beam search mixes information across
the batch dimension -- SURVEY.md section 8e), so rank r captions a contiguous slice of the batch with
its own engine; the only exchange is one final all-gather of the caption ids (and log-probs).
"""

from __future__ import annotations

import os
from datetime import timedelta
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None, timeout_s: float = 120.0) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group
    when WORLD_SIZE > 1 (NCCL on GPUs, gloo otherwise).  `timeout_s` bounds every collective: a rank whose
    peer died raises after that long instead of waiting in a barrier until somebody kills the job."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # bind the communicator to this rank's GPU and create it now (not lazily at the first collective,
            # which would otherwise happen in the middle of the pipelined work)
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=timedelta(seconds=timeout_s),
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world, timeout=timedelta(seconds=timeout_s))
    return rank, world, local_rank


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `total` images for `rank`; the first total % world ranks get one more."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_tensor(t: Optional[torch.Tensor], world: int, rank: int) -> Optional[torch.Tensor]:
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def gather_captions(ids: torch.Tensor, logp: torch.Tensor, total: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the per-rank (b_r, ..., T) ids / log-probs into the full (total, ..., T) batch order.

    Shards may differ by one image, so every rank pads to the largest shard before the collective
    (NCCL all_gather needs equal sizes) and the padding is dropped afterwards.  ids travel as int32.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ids, logp
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    biggest = max(sizes)
    tail = tuple(ids.shape[1:])

    def padded(x, dtype):
        out = torch.zeros((biggest,) + tail, dtype=dtype, device=x.device)
        out[: x.shape[0]] = x.to(dtype)
        return out

    ids32, lp = padded(ids, torch.int32), padded(logp, torch.float32)
    all_ids = [torch.empty_like(ids32) for _ in range(world)]
    all_lp = [torch.empty_like(lp) for _ in range(world)]
    dist.all_gather(all_ids, ids32, group=group)
    dist.all_gather(all_lp, lp, group=group)
    full_ids = torch.cat([a[:s] for a, s in zip(all_ids, sizes)], 0).to(torch.int64)
    full_lp = torch.cat([a[:s] for a, s in zip(all_lp, sizes)], 0)
    return full_ids, full_lp


# ---------------------------------------------------------------------------------------------------------------
# Data-parallel training (the training step's only exchange): every rank runs the step on its shard of the batch; the
# loss is the mean over the valid target tokens of the WHOLE batch, so the ranks first agree on that count, weight
# their tokens by 1 / count, and sum the flat gradient buffers.  The result is the single-process gradient of the
# global batch (up to fp32 summation order), not an average of per-rank means.
# ---------------------------------------------------------------------------------------------------------------
def global_token_weight(n_valid_local: torch.Tensor, group=None) -> torch.Tensor:
    """1 / (number of valid target tokens over all ranks), a 0-d fp32 tensor on n_valid_local's device (no host sync)."""
    total = n_valid_local.detach().to(torch.float32).clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / total


def sum_gradients_(flat_grads: torch.Tensor, loss: Optional[torch.Tensor] = None, group=None) -> None:
    """In-place sum over ranks of the flat gradient buffer (ONE collective) and of the loss contribution."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    if loss is not None:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
