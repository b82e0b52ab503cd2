"""CPU-only measurement of the host glue either side of the caption path:
native collate / caption decode against the reference's Python procedure, on the bench workload's shapes
(256 images x 49 rows x 2048 fp32 features in; 256 x 20 ids out, V 10201).  Prints one JSON line.

The "reference procedure" legs restate what the reference executes (utils/instance.py:42-49,156-171 followed by
the bf16 cast the engine needs; data_utils/vocab.py:104-122 + trainers/vi_trainer.py:251); they are the baseline,
not product code.

usage:  python tools/bench_host_glue.py [--batch 256] [--rows 49] [--width 2048] [--repeat 20]
"""

from __future__ import annotations

import argparse
import itertools
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from openviic_b200.data_utils import FeatureBatcher, Vocab  # noqa: E402


def best_of(fn, repeat):
    fn()
    times = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return min(times)


def reference_collate(rows):
    padded, longest = [], max(r.shape[0] for r in rows)
    for r in rows:
        t = torch.tensor(r)
        if t.shape[0] < longest:
            t = torch.cat([t, torch.zeros((longest - t.shape[0], t.shape[-1]))], dim=0)
        padded.append(t.unsqueeze(0))
    return torch.cat(padded, dim=0).to(torch.bfloat16)


def reference_decode(vocab, ids):
    out = []
    for vec in ids:
        words = []
        for idx in vec.tolist():
            if vocab.itos[idx] not in vocab.specials:
                words.append(vocab.itos[idx])
            if idx == vocab.eos_idx:
                break
        out.append(" ".join(k for k, _ in itertools.groupby(" ".join(words).strip().split())))
    return out


def cider_leg(rng, threads, repeat):
    """The self-critical reward of one step: 64 images x beam 5 hypotheses, 5 references each, document frequencies
    from a 10 000-image corpus.  The reference leg runs the reference's own evaluation/cider when /root/reference
    (or $OPENVIIC_REFERENCE) exists -- build container only."""
    from openviic_b200.evaluation import Cider
    words = [f"w{i}" for i in range(3000)]
    p = 1.0 / np.arange(1, len(words) + 1)
    p /= p.sum()
    corpus = {str(i): [" ".join(rng.choice(words, size=rng.integers(6, 18), p=p)) for _ in range(5)] for i in range(10000)}
    images, beam = 64, 5
    gts = {str(i): corpus[str(i // beam)] for i in range(images * beam)}
    res = {str(i): [" ".join(rng.choice(words, size=rng.integers(6, 18), p=p))] for i in range(images * beam)}
    ours = Cider(corpus, threads=threads)
    got = ours.compute_score(gts, res)[1]
    t_native = best_of(lambda: ours.compute_score(gts, res), repeat)
    out = {"workload": "320 hypotheses x 5 references, 10 000-image corpus", "native_hypotheses_per_s": len(res) / t_native}
    ref_root = Path(os.environ.get("OPENVIIC_REFERENCE", "/root/reference"))
    if ref_root.exists():
        sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle" / "ref_harness" / "shims"))
        sys.path.insert(0, str(ref_root))
        from evaluation.cider import Cider as RefCider
        ref = RefCider(corpus)
        want = ref.compute_score(gts, res)[1]
        assert np.abs(want - got).max() <= 1e-12
        t_ref = best_of(lambda: ref.compute_score(gts, res), 3)
        out.update({"reference_hypotheses_per_s": len(res) / t_ref, "speedup": t_ref / t_native})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--rows", type=int, default=49)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--repeat", type=int, default=20)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    threads = args.threads or min(16, os.cpu_count() or 1)
    torch.set_num_threads(threads)
    rng = np.random.default_rng(0)
    lengths = rng.integers(args.rows // 2, args.rows + 1, size=args.batch)
    lengths[0] = args.rows
    rows = [rng.standard_normal((n, args.width)).astype(np.float32) for n in lengths]
    batcher = FeatureBatcher(args.batch, args.rows, args.width, threads=threads)
    got, _ = batcher.collate(rows)
    assert torch.equal(got, reference_collate(rows))
    t_native = best_of(lambda: batcher.collate(rows), args.repeat)
    t_ref = best_of(lambda: reference_collate(rows), max(3, args.repeat // 4))

    vocab = Vocab.from_itos(["<pad>", "<bos>", "<eos>", "<unk>"] + [f"w{i}" for i in range(4, 10201)], 20)
    ids = torch.from_numpy(rng.integers(4, 10201, size=(args.batch, 20)))
    ids[:, 0] = vocab.bos_idx
    for r in range(args.batch):   # caption lengths 8..19, as beam search leaves them: eos, then padding
        stop = rng.integers(8, 20)
        ids[r, stop] = vocab.eos_idx
        ids[r, stop + 1:] = vocab.padding_idx
    assert vocab.decode_predictions(ids) == reference_decode(vocab, ids)
    d_native = best_of(lambda: vocab.decode_predictions(ids), args.repeat * 5)
    d_ref = best_of(lambda: reference_decode(vocab, ids), args.repeat)

    cider = cider_leg(rng, threads, args.repeat)

    in_bytes = int(sum(r.nbytes for r in rows))
    print(json.dumps({
        "workload": f"{args.batch} images x <= {args.rows} rows x {args.width} fp32 -> bf16 batch; {args.batch} x 20 ids -> text, V 10201",
        "host_threads": threads,
        "collate": {"native_images_per_s": args.batch / t_native, "reference_procedure_images_per_s": args.batch / t_ref,
                    "native_gb_per_s_read": in_bytes / t_native / 1e9, "speedup": t_ref / t_native},
        "decode": {"native_captions_per_s": args.batch / d_native, "reference_procedure_captions_per_s": args.batch / d_ref,
                   "speedup": d_ref / d_native},
        "cider": cider,
    }))


if __name__ == "__main__":
    main()
