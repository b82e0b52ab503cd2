"""Generate tests/golden/cider.json by running the REAL reference CIDEr (evaluation/cider) on CPU.

Test infrastructure (build container only).  Pins SURVEY.md section 8f row 3: the self-critical reward /
evaluation CIDEr-D -- reference evaluation/cider/cider.py:12-38 and cider_scorer.py:9-167, called as
trainers/vi_trainer.py:34 (``Cider(train captions)``) and :137-145 (``compute_score(gts, gens)[1]``).
The fixture stores the captions (inputs) and the reference's scores (outputs) for

  * corpus mode: document frequencies from a 300-image training corpus, a 5-beam batch scored against the
    images' own references, each image's references repeated per beam as the trainer does;
  * batch mode: ``Cider()`` without a corpus (document frequencies from the batch's references);
  * edge cases: empty hypothesis, hypothesis equal to a reference, one-word captions, words unseen in the
    corpus, a single reference, repeated words (clipping), very long hypothesis (length penalty).

usage:  python oracle/ref_harness/gen_golden_cider.py
"""

from __future__ import annotations

import itertools
import json
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path(os.environ.get("OPENVIIC_REFERENCE", "/root/reference"))
sys.path.insert(0, str(HERE / "shims"))
sys.path.insert(0, str(REFERENCE))

from evaluation.cider import Cider as RefCider  # noqa: E402

WORDS = ("một người đàn ông phụ nữ đang đi bộ trên đường phố bán hàng ở chợ ngồi ghế trước cửa xe máy màu đỏ đậu "
         "hai những chiếc có nhiều xung quanh biển hiệu cà phê trà treo mới mở đông đúc lái buýt mua").split()


def main() -> None:
    rng = np.random.default_rng(2024)
    weights = 1.0 / np.arange(1, len(WORDS) + 1)          # a Zipf-like word distribution: shared n-grams do occur
    weights /= weights.sum()

    def caption(lo=3, hi=16):
        return " ".join(rng.choice(WORDS, size=rng.integers(lo, hi), p=weights))

    corpus = {f"{i}": [caption() for _ in range(rng.integers(1, 6))] for i in range(300)}
    beam, images = 5, 24
    picked = rng.choice(300, size=images, replace=False)
    caps_gt = [corpus[f"{i}"] for i in picked]
    caps_gt = list(itertools.chain(*([a, ] * beam for a in caps_gt)))      # trainers/vi_trainer.py:140
    caps_gen = []
    for refs in caps_gt:                                                    # hypotheses near the references
        words = refs[rng.integers(len(refs))].split()
        keep = [w for w in words if rng.random() > 0.25] + list(rng.choice(WORDS, size=rng.integers(0, 3)))
        caps_gen.append(" ".join(keep))
    caps_gen[0] = ""                                   # empty hypothesis
    caps_gen[1] = caps_gt[1][0]                        # identical to a reference
    caps_gen[2] = "một"                                # one word
    caps_gen[3] = "từ lạ chưa gặp bao giờ"             # only unseen words
    caps_gen[4] = "một một một một người người"         # repeats: clipping
    caps_gen[5] = " ".join(["một người đàn ông đang đi bộ"] * 6)   # long: length penalty
    caps_gen[6] = "  một   người  đang đi  "           # stray whitespace
    gens = {f"{i}": [c, ] for i, c in enumerate(caps_gen)}
    gts = {f"{i}": c for i, c in enumerate(caps_gt)}

    with_corpus = RefCider(corpus).compute_score(gts, gens)
    batch_only = RefCider().compute_score(gts, gens)
    single = RefCider(corpus).compute_score({"a": ["một người đang đi bộ"]}, {"a": ["một người đi bộ"]})
    other_sigma = RefCider(corpus, sigma=3.0).compute_score(gts, gens)   # n != 4 raises in the reference: cook_refs ignores n

    golden = {
        "corpus": corpus, "gts": gts, "gens": gens, "beam": beam,
        "with_corpus": {"mean": float(with_corpus[0]), "scores": [float(s) for s in with_corpus[1]]},
        "batch_only": {"mean": float(batch_only[0]), "scores": [float(s) for s in batch_only[1]]},
        "single": {"mean": float(single[0]), "scores": [float(s) for s in single[1]]},
        "sigma3": {"mean": float(other_sigma[0]), "scores": [float(s) for s in other_sigma[1]]},
    }
    out = REPO / "tests" / "golden" / "cider.json"
    with open(out, "w", encoding="utf-8") as fh:
        json.dump(golden, fh, ensure_ascii=False, indent=1)
    print(f"wrote {out}: {len(gts)} hypotheses, corpus mean {with_corpus[0]:.4f}, batch-only mean {batch_only[0]:.4f}")


if __name__ == "__main__":
    main()
