"""One batch of the bench workload through the engine, eagerly (for ncu captures)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from openviic_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
cfg, vocab, model, weights = bench.build_model("standard_grid", dev)
eng = model.engine(B, 49, 5)
feats = synthetic.synth_features(B, 49, 2048, 1234, ragged=False).to(torch.bfloat16).to(dev)
for _ in range(2):
    eng.encode(feats, None)
    ids, lp = eng.beam_search(out_size=1, use_graph=False)
torch.cuda.synchronize()
print("ok", ids.shape)
