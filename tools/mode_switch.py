"""Host-buffer and device-buffer entry points in alternation (regression check for a hang seen when switching)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from openviic_b200 import CaptionEngine, synthetic
dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg, vocab, model, weights = bench.build_model("standard_grid", dev)
engs = []
for _ in range(S):
    e = CaptionEngine(cfg.MODEL, vocab, model.state_dict(), dev); e.reserve(256, 49, 5); engs.append(e)
streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
fh = [synthetic.synth_features(256, 49, 2048, 1234 + i, ragged=False).to(torch.bfloat16).pin_memory() for i in range(4)]
fd = [f.to(dev) for f in fh]
outs_d = [(torch.empty((256, 1, 20), device=dev, dtype=torch.int64), torch.empty((256, 1, 20), device=dev, dtype=torch.float32)) for _ in range(S)]
outs_h = [(torch.empty((256, 1, 20), dtype=torch.int64).pin_memory(), torch.empty((256, 1, 20), dtype=torch.float32).pin_memory()) for _ in range(S)]
def run(mode, steps):
    for i in range(steps):
        k = i % S
        with torch.cuda.stream(streams[k]):
            if mode == "host":
                engs[k].caption_host(fh[i % 4], None, 1, True, outs_h[k], sync=False)
            else:
                engs[k].caption_device(fd[i % 4], None, 1, True, outs_d[k])
for mode in ["host", "dev", "host", "dev"]:
    run(mode, 3 * S); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(mode, STEPS); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(mode, round(STEPS * 256 / dt), flush=True)
print("ids equal:", bool((outs_h[0][0] == outs_d[0][0].cpu()).all()))
