"""Two XE training steps at the bench shape, eagerly (for ncu launch lists)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from openviic_b200 import synthetic  # noqa: E402
from openviic_b200.training import XETrainer  # noqa: E402

dev = torch.device("cuda:0")
cfg, vocab, model, weights = bench.build_model("standard_grid", dev)
trainer = XETrainer(model, lr=1.0, warmup=10000, dropout_seed=1)
f = synthetic.synth_features(256, 49, 2048, 77, ragged=False).to(torch.bfloat16).to(dev)
tok, tgt = synthetic.synth_captions(256, bench.MAX_LEN, bench.VOCAB, 77)
for _ in range(2):
    loss = trainer.step(f, tok.to(dev), tgt.to(dev))
torch.cuda.synchronize()
print("ok", loss.item())
