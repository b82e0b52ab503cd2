"""Parity cases shared by the golden-fixture generator and the tests (test infrastructure).

Every case is fully determined by this table: YAML config (+ overrides), synthetic-input shape and
seed.  ``std_region_A`` is BASELINE.json's config A at full size (the reference's own CPU-runnable
case); the others are the remaining configs at sizes the oracle finishes in seconds.
"""

CASES = {
    # name: config file, batch, visual tokens, beam, max_len (T), vocab, seed [, overrides]
    "std_region_A": dict(config="standard_transformer_using_region.yaml", batch=16, n=50, beam=3, max_len=20,
                         vocab=10201, seed=11),
    "std_grid": dict(config="standard_transformer.yaml", batch=6, n=49, beam=5, max_len=20, vocab=1000, seed=12),
    "m2": dict(config="meshed_memory_transformer.yaml", batch=5, n=50, beam=5, max_len=20, vocab=1000, seed=13),
    "ort": dict(config="object_relation_transformer.yaml", batch=5, n=50, beam=5, max_len=20, vocab=1000, seed=14),
    "ort_trig": dict(config="object_relation_transformer.yaml", batch=4, n=50, beam=5, max_len=12, vocab=600, seed=15,
                     overrides={"MODEL": {"ENCODER": {"TRIGNOMETRIC_EMBEDDING": True}}}),
    "aoa": dict(config="attention_on_attention.yaml", batch=4, n=50, beam=5, max_len=16, vocab=800, seed=16),
    "aug_mem": dict(config="augmented_memory_transformer.yaml", batch=4, n=37, beam=4, max_len=16, vocab=777, seed=17),
    # CamoTransformer: CrossAttentionMultiLevelEncoder (one 64-wide head in the encoder); module-level CUDA path
    "camo": dict(config="camo_transformer.yaml", batch=4, n=44, beam=5, max_len=14, vocab=900, seed=18),
    # dual-path DLCT encoder (n regions + grid x grid cells), standard decoder; module-level CUDA path; the reference side
    # is the patched composition of oracle/ref_harness/gen_golden_dlct.py (the reference has no such architecture)
    "dlct": dict(config="dlct_transformer.yaml", batch=4, n=30, grid=7, beam=5, max_len=12, vocab=800, seed=19,
                 overrides={"MODEL": {"ENCODER": {"LAYERS": 2}}}),
}


def _merge(node, overrides):
    for key, value in overrides.items():
        if isinstance(value, dict):
            _merge(node[key], value)
        else:
            node[key] = value


def apply_overrides(cfg, case):
    """Apply a case's nested overrides to a config node (works for our CfgNode and the yacs shim)."""
    if case.get("overrides"):
        _merge(cfg, case["overrides"])
    return cfg


# cases the whole-path engine covers (the others run on the registered modules: the module-level CUDA path)
MODULE_PATH_CASES = ("camo", "dlct")
ENGINE_CASES = tuple(c for c in CASES if c not in MODULE_PATH_CASES)

# operator-level case (no runnable architecture reaches this class in the reference): oracle/ref_harness/gen_golden_ops.py
ADAPTIVE_ATTENTION_CASE = dict(
    config=dict(ARCHITECTURE="AdaptiveScaledDotProductAttention", D_MODEL=512, HEAD=8, D_KEY=64, D_VALUE=64, DROPOUT=0.1),
    batch=5, nq=7, nk=37, seed=21)

# T1, the XE training step (oracle/ref_harness/gen_golden_train.py): `steps` optimizer steps on fresh synthetic batches
TRAIN_CASES = {
    "std_grid": dict(config="standard_transformer.yaml", batch=8, n=49, max_len=20, vocab=1000, seed=31, steps=3,
                     lr=1.0, warmup=10000),
    "std_region": dict(config="standard_transformer_using_region.yaml", batch=8, n=50, max_len=16, vocab=777, seed=32,
                       steps=2, lr=1.0, warmup=100, dropout_seed=9001, scst_beam=5, rl_lr=5e-6, eos_scale=4.0),
}

# T1 on the other architectures: the ORACLE's training step pinned to the reference (no GPU trainer for them yet -- the
# backward kernels cover the standard transformer; these fixtures are the checker the next ones will be held to)
ORACLE_ONLY_TRAIN_CASES = {
    "m2": dict(config="meshed_memory_transformer.yaml", batch=4, n=50, max_len=12, vocab=600, seed=41, steps=2, lr=1.0, warmup=100),
    "ort": dict(config="object_relation_transformer.yaml", batch=4, n=50, max_len=12, vocab=600, seed=42, steps=2, lr=1.0, warmup=100),
    "aoa": dict(config="attention_on_attention.yaml", batch=4, n=50, max_len=12, vocab=600, seed=43, steps=2, lr=1.0, warmup=100),
}
