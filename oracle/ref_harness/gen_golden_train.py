"""Golden fixture of the XE training step (T1): the REAL reference model, loss and optimizer.

Builds the reference model from this repo's YAML, loads the synthetic weights, and runs the body of
trainers/vi_trainer.py:105-119 twice: with every nn.Dropout set to p = 0, and (cases with "dropout_seed") with every
nn.Dropout module's forward WRAPPED to apply the counter-based mask of oracle.dropout_keep at the site named by the
module's qualified name (torch's own dropout generator has no portable restatement; p, the sites and the scaling are the
reference's) --
out = model(items); loss = NLLLoss(ignore_index=pad)(out.view(-1, V), shifted.view(-1)); backward; Adam(lr, betas=(0.9,
0.98)).step(); LambdaLR(lambda_lr).step() (trainers/base_trainer.py:89-91, 114-117) -- for STEPS batches.  It refuses
to write the fixture unless oracle.xe_train_steps reproduces losses, first-step gradients and final weights.

usage:  python oracle/ref_harness/gen_golden_train.py
"""

from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch
from torch.nn import NLLLoss
from torch.optim import Adam
from torch.optim.lr_scheduler import LambdaLR

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path(os.environ.get("OPENVIIC_REFERENCE", "/root/reference"))
sys.path.insert(0, str(HERE / "shims"))
sys.path.insert(0, str(REFERENCE))

import models  # noqa: E402,F401  (reference package: fills the registries)
from builders.model_builder import build_model as ref_build_model  # noqa: E402
from configs.utils import get_config as ref_get_config  # noqa: E402
from utils.instance import InstanceList as RefInstanceList  # noqa: E402

sys.path.insert(0, str(REPO))
from openviic_b200 import synthetic  # noqa: E402
from openviic_b200.configs import get_config  # noqa: E402
from oracle import caption_oracle as oracle  # noqa: E402
from oracle.cases import ORACLE_ONLY_TRAIN_CASES, TRAIN_CASES, apply_overrides  # noqa: E402

sys.path.insert(0, str(HERE))
import gen_golden  # noqa: E402,F401  (applies the documented ORT encoder_forward call-site patch to the reference class)


def run_case(name: str, case: dict, dropout: bool = False) -> dict:
    cfg_path = REPO / "openviic_b200" / "configs" / case["config"]
    ref_cfg = apply_overrides(ref_get_config(str(cfg_path)), case)
    ref_cfg.MODEL.DEVICE = "cpu"
    vocab = synthetic.SyntheticVocab(case["vocab"], case["max_len"])
    torch.manual_seed(0)
    model = ref_build_model(ref_cfg.MODEL, vocab)
    weights = synthetic.load_synthetic_weights(model, case["seed"])
    model.train()
    seeds = None
    if dropout:
        seeds = [case["dropout_seed"] + i for i in range(case["steps"])]
        step_box = {"i": 0}
        for qual, m in model.named_modules():
            if isinstance(m, torch.nn.Dropout) and m.p > 0:
                def masked(x, _name=qual, _p=m.p):
                    keep = oracle.dropout_keep(x.numel(), seeds[step_box["i"]], _name, _p).view(x.shape)
                    return x * keep * torch.tensor(1.0 / (1.0 - float(_p)), dtype=torch.float32)
                m.forward = masked
    else:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    batches = synthetic.synth_train_batches(ref_cfg.MODEL, case)
    optim = Adam(model.parameters(), lr=case["lr"], betas=(0.9, 0.98))
    d_model, warmup = model.encoder.d_model, case["warmup"]
    sched = LambdaLR(optim, lambda step: (d_model ** -.5) * min((step + 1) ** -.5, (step + 1) * warmup ** -1.5))
    loss_fn = NLLLoss(ignore_index=vocab.padding_idx)
    losses, first_grads = [], None
    for field, feats, tokens, targets, boxes in batches:
        items = RefInstanceList()
        items.set(field, feats)
        if boxes is not None:
            items.set("region_boxes", boxes)
        items.set("caption_tokens", tokens)
        out = model(items).contiguous()
        optim.zero_grad()
        loss = loss_fn(out.view(-1, len(vocab)), targets.view(-1))
        loss.backward()
        if first_grads is None:
            first_grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
        optim.step()
        sched.step()
        losses.append(float(loss.detach()))
        if dropout:
            step_box["i"] += 1
    final = {k: v.detach().clone() for k, v in model.state_dict().items()}

    our_cfg = apply_overrides(get_config(cfg_path), case)
    o_final, o_losses, o_grads = oracle.xe_train_steps(
        weights, our_cfg.MODEL, vocab, [(f, t, y, b) for _, f, t, y, b in batches], case["lr"], case["warmup"], dropout_seeds=seeds)
    worst_g = max(float((first_grads[k] - o_grads[k]).abs().max()) for k in first_grads if first_grads[k] is not None)
    missing = [k for k in first_grads if (first_grads[k] is None) != (o_grads.get(k) is None)]
    worst_w = max(float((final[k].float() - o_final[k]).abs().max()) for k in final if k in o_final and final[k].numel())
    worst_l = max(abs(a - b) for a, b in zip(losses, o_losses))
    print(f"[{name}{' + dropout' if dropout else ''}] reference losses {losses}; oracle-vs-reference: loss {worst_l:.2e}, first-step gradients {worst_g:.2e}, "
          f"weights after {len(batches)} steps {worst_w:.2e}, parameters whose gradient exists on one side only: {missing}")
    # exact (0.0) on the standard transformer; the meshed / geometric / AoA graphs sum a few terms in another order, which
    # shows as fp32 rounding in loss and gradients -- and, through Adam's sign-like first steps (a gradient entry that is
    # zero up to rounding moves its weight by +-lr), as differences of up to 2 lr per step in a handful of weights
    lr_max = case["lr"] * (d_model ** -.5) * case["steps"] * warmup ** -1.5
    if worst_l > 1e-5 or worst_g > 1e-6 or worst_w > max(1e-6, 2.2 * case["steps"] * lr_max) or missing:
        raise SystemExit(f"oracle does not reproduce the reference's training step on case {name}")

    # the fixture: losses, and for every parameter the first-step gradient's and the final weight's norm plus a strided
    # sample (the full tensors are regenerated by the oracle in the tests; this pins the oracle to the reference)
    fixture = {"losses": np.asarray(losses, dtype=np.float64)}
    for k, g in first_grads.items():
        if g is None:
            continue
        flat = g.reshape(-1)
        fixture["g/" + k] = np.concatenate([[float(flat.norm())], flat[:: max(1, flat.numel() // 16)][:16].numpy()])
        wf = final[k].reshape(-1)
        fixture["w/" + k] = np.concatenate([[float(wf.norm())], wf[:: max(1, wf.numel() // 16)][:16].numpy()])
    return fixture


def run_scst(name: str, case: dict) -> dict:
    """The self-critical step: the REAL reference's beam search with autograd on, its loss expression and Adam (dropout
    p = 0; rewards are synthetic constants -- the CIDEr scorer is pinned separately by gen_golden_cider.py)."""
    cfg_path = REPO / "openviic_b200" / "configs" / case["config"]
    ref_cfg = apply_overrides(ref_get_config(str(cfg_path)), case)
    ref_cfg.MODEL.DEVICE = "cpu"
    vocab = synthetic.SyntheticVocab(case["vocab"], case["max_len"])
    torch.manual_seed(0)
    model = ref_build_model(ref_cfg.MODEL, vocab)
    weights = synthetic.load_synthetic_weights(model, case["seed"])
    synthetic.boost_eos(model, weights, vocab.eos_idx, case["eos_scale"])
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    field, feats, tokens, targets, boxes = synthetic.synth_train_batches(ref_cfg.MODEL, case)[0]
    b, beam = case["batch"], case["scst_beam"]
    items = RefInstanceList()
    items.set(field, feats)
    if boxes is not None:
        items.set("region_boxes", boxes)
    optim = Adam(model.parameters(), lr=case["rl_lr"])
    outs, log_probs = model.beam_search(items, batch_size=b, beam_size=beam, out_size=beam)      # vi_trainer.py:132-133
    reward = synthetic.synth_rewards(b, beam, case["seed"])
    optim.zero_grad()
    reward_baseline = torch.mean(reward, dim=-1, keepdim=True)
    loss = -torch.mean(log_probs, -1) * (reward - reward_baseline)                                 # :146
    loss = loss.mean()
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in model.named_parameters()}
    optim.step()
    final = {k: v.detach().clone() for k, v in model.state_dict().items()}

    our_cfg = apply_overrides(get_config(cfg_path), case)
    captions = outs.detach()
    with torch.no_grad():
        _, o_tok_lp = oracle.scst_loss(weights, our_cfg.MODEL, vocab, feats, captions, reward, boxes)
    o_final, o_loss, o_grads = oracle.scst_step(weights, our_cfg.MODEL, vocab, feats, captions, reward, case["rl_lr"], boxes)
    worst_lp = float((log_probs.detach() - o_tok_lp).abs().max())
    # key-projection biases have a mathematically zero gradient (a key bias shifts all logits of a query alike): both sides
    # hold rounding noise there (1e-10), compared absolutely; every other parameter relative to its largest entry
    worst_g = max(float((grads[k] - o_grads[k]).abs().max() / grads[k].abs().max().clamp_min(1e-12))
                  for k in grads if grads[k] is not None and not k.endswith("fc_k.bias"))
    worst_zero = max(float(max(grads[k].abs().max(), o_grads[k].abs().max())) for k in grads if k.endswith("fc_k.bias"))
    missing = [k for k in grads if (grads[k] is None) != (o_grads.get(k) is None)]
    worst_w = max(float((final[k].float() - o_final[k]).abs().max()) for k in final if k in o_final and final[k].numel())
    n_eos = int((captions == vocab.eos_idx).sum())
    print(f"[{name} scst] reference loss {float(loss.detach()):.8f}, captions with <eos>: {n_eos}; oracle (teacher-forced restatement) vs reference "
          f"(backward through the beam search): loss {abs(float(loss) - o_loss):.2e}, per-token log-probs {worst_lp:.2e}, gradients "
          f"(relative to each parameter's largest entry) {worst_g:.2e} (key biases, zero by construction: {worst_zero:.1e}), weights after the Adam step {worst_w:.2e}, one-sided: {missing}")
    if abs(float(loss) - o_loss) > 1e-6 or worst_lp > 1e-4 or worst_g > 1e-4 or worst_zero > 1e-7 or missing:
        raise SystemExit(f"the teacher-forced restatement of the self-critical step does not reproduce the reference on {name}")
    fixture = {"loss": np.float64(float(loss)), "captions": captions.numpy(), "rewards": reward.numpy(), "log_probs": log_probs.detach().numpy()}
    for k, g in grads.items():
        if g is None:
            continue
        flat = g.reshape(-1)
        fixture["g/" + k] = np.concatenate([[float(flat.norm())], flat[:: max(1, flat.numel() // 16)][:16].numpy()])
    return fixture


def main():
    for name, case in ORACLE_ONLY_TRAIN_CASES.items():
        out = REPO / "tests" / "golden" / f"train_{name}.npz"
        np.savez_compressed(out, **run_case(name, case))
        print("wrote", out, out.stat().st_size, "bytes")
    for name, case in TRAIN_CASES.items():
        if "scst_beam" in case:
            out = REPO / "tests" / "golden" / f"train_{name}_scst.npz"
            np.savez_compressed(out, **run_scst(name, case))
            print("wrote", out, out.stat().st_size, "bytes")
        out = REPO / "tests" / "golden" / f"train_{name}.npz"
        np.savez_compressed(out, **run_case(name, case))
        print("wrote", out, out.stat().st_size, "bytes")
        if "dropout_seed" in case:
            out = REPO / "tests" / "golden" / f"train_{name}_dropout.npz"
            np.savez_compressed(out, **run_case(name, case, dropout=True))
            print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
