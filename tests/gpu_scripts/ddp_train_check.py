"""GPU probe (not collected by pytest; needs 2 GPUs): the data-parallel XE training step against the single-process step
on the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29640 \
        tests/gpu_scripts/ddp_train_check.py
"""
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import openviic_b200 as ov  # noqa: E402
from openviic_b200 import parallel, synthetic  # noqa: E402
from openviic_b200.training import XETrainer  # noqa: E402
from oracle.cases import TRAIN_CASES, apply_overrides  # noqa: E402  (case table only)


def main():
    rank, world, local = parallel.init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    case = TRAIN_CASES["std_region"]
    cfg = apply_overrides(ov.get_config(case["config"]), case)
    cfg.MODEL.DEVICE = str(dev)
    vocab = synthetic.SyntheticVocab(case["vocab"], case["max_len"])

    def fresh():
        model = ov.build_model(cfg.MODEL, vocab).to(dev)
        synthetic.load_synthetic_weights(model, case["seed"])
        return XETrainer(model, lr=case["lr"], warmup=case["warmup"], dropout_seed=None, ignore_dropout=True)

    batches = synthetic.synth_train_batches(cfg.MODEL, case)
    trainer = fresh()
    losses = []
    for _, feats, tokens, targets, _ in batches:
        lo, hi = parallel.shard_bounds(feats.shape[0], world, rank)
        losses.append(trainer.step(feats[lo:hi].to(dev).to(torch.bfloat16), tokens[lo:hi].to(dev), targets[lo:hi].to(dev)))
    torch.cuda.synchronize()
    if rank == 0:
        single = fresh()
        ref_losses = []
        with torch.no_grad():
            for _, feats, tokens, targets, _ in batches:
                ref_losses.append(single.loss_and_grads(feats.to(dev).to(torch.bfloat16), tokens.to(dev), targets.to(dev)))
                single.optimizer_step()
        torch.cuda.synchronize()
        start = fresh().p32
        du = trainer.p32 - start
        dr = single.p32 - start
        cos = (du * dr).sum().item() / math.sqrt((du * du).sum().item() * (dr * dr).sum().item())
        out = {"world": world, "losses": [x.item() for x in losses], "single_process_losses": [x.item() for x in ref_losses],
               "update_cosine": cos, "max_abs_param_diff": (trainer.p32 - single.p32).abs().max().item(),
               "max_abs_update": dr.abs().max().item()}
        print(json.dumps(out))
        assert all(abs(a - b) < 2e-3 for a, b in zip(out["losses"], out["single_process_losses"])) and cos > 0.99
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
