"""tests/golden/state_dict_keys.json: parameter / buffer names and shapes of the REAL reference models, one entry per
YAML file shipped in openviic_b200/configs/ (vocabulary of 300 words, captions of 20 tokens).  The dual-path entry is
the composition of reference modules that gen_golden_dlct.py builds (the reference has no architecture for it).

usage:  python oracle/ref_harness/gen_golden_keys.py
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import gen_golden_dlct as dlct  # noqa: E402  (sets up the import path of the reference and the shims)
from builders.model_builder import build_model as ref_build_model  # noqa: E402
from configs.utils import get_config as ref_get_config  # noqa: E402

from openviic_b200 import synthetic  # noqa: E402

REPO = dlct.REPO


def main():
    vocab = synthetic.SyntheticVocab(300, 20)
    table = {}
    for path in sorted((REPO / "openviic_b200" / "configs").glob("*.yaml")):
        cfg = ref_get_config(str(path))
        cfg.MODEL.DEVICE = "cpu"
        if path.name == "dlct_transformer.yaml":
            model = dlct.DLCTReference(cfg.MODEL, vocab)
        else:
            model = ref_build_model(cfg.MODEL, vocab)
        table[path.name] = {k: list(v.shape) for k, v in model.state_dict().items()}
        print(f"{path.name}: {len(table[path.name])} entries")
    out = REPO / "tests" / "golden" / "state_dict_keys.json"
    out.write_text(json.dumps(table, indent=0, sort_keys=True) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
