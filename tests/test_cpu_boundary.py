"""Host-side boundary: registry, builders, config loader, inputs, C-ABI surface (no GPU needed)."""

import json
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

import openviic_b200 as ov
from openviic_b200 import cabi, engine as engine_mod, synthetic
from openviic_b200.builders.registry import Registry

REPO = Path(__file__).resolve().parent.parent
CONFIGS = sorted(p.name for p in (REPO / "openviic_b200" / "configs").glob("*.yaml"))


def test_registry_contract():
    reg = Registry("DEMO")

    @reg.register()
    class Foo:
        pass

    class Bar:
        pass

    reg.register(Bar)
    assert reg.get("Foo") is Foo and reg.get("Bar") is Bar and "Foo" in reg and len(reg) == 2
    assert dict(iter(reg)) == {"Foo": Foo, "Bar": Bar}
    with pytest.raises(KeyError, match="No object named 'Nope' found in 'DEMO' registry"):
        reg.get("Nope")
    with pytest.raises(AssertionError, match="already registered"):
        reg.register(Foo)
    assert "Foo" in repr(reg)


def test_every_yaml_name_resolves():
    expected = {
        ov.META_ATTENTION: ["ScaledDotProductAttention", "AugmentedGeometryScaledDotProductAttention",
                            "AugmentedMemoryScaledDotProductAttention"],
        ov.META_ENCODER: ["Encoder", "MultilevelEncoder", "GeometricEncoder"],
        ov.META_DECODER: ["Decoder", "MeshedDecoder"],
        ov.META_ARCHITECTURE: ["StandardTransformerUsingRegion", "StandardTransformerUsingGrid",
                               "MeshedMemoryTransformer", "ObjectRelationTransformer"],
        ov.META_VISION_EMBEDDING: ["FeatureEmbedding"],
        ov.META_TEXT_EMBEDDING: ["UsualEmbedding"],
    }
    for reg, names in expected.items():
        for name in names:
            assert name in reg, name


@pytest.mark.parametrize("name", CONFIGS)
def test_models_build_with_reference_state_dict_layout(name):
    cfg = ov.get_config(name)
    cfg.MODEL.DEVICE = "cpu"
    vocab = synthetic.SyntheticVocab(300, 20)
    model = ov.build_model(cfg.MODEL, vocab)
    ours = {k: list(v.shape) for k, v in model.state_dict().items()}
    ref = json.load(open(REPO / "tests" / "golden" / "state_dict_keys.json"))[name]   # dumped from the reference
    assert ours == ref
    if name in ("dlct_transformer.yaml", "camo_transformer.yaml"):   # module-level CUDA path only
        assert not model.engine_supported()
        return
    assert model.engine_supported()
    desc = engine_mod.model_desc(cfg.MODEL, vocab)
    assert (desc.d_model, desc.heads, desc.d_k, desc.vocab, desc.max_len) == (512, 8, 64, 300, 20)
    assert desc.n_enc_levels == (3 if cfg.MODEL.DECODER.ARCHITECTURE == "MeshedDecoder" else 1)


def test_config_loader_reads_reference_yaml_schema(tmp_path):
    cfg = ov.get_config("meshed_memory_transformer.yaml")
    assert cfg.MODEL.ENCODER.SELF_ATTENTION.MEMORY == 40 and cfg.MODEL.DECODER.ATTENTION.N_ENCODER_LAYERS == 3
    assert cfg.MODEL.DECODER.TEXT_EMBEDDING.WORD_EMBEDDING is None
    clone = cfg.clone()
    clone.MODEL.DEVICE = "cpu"
    assert cfg.MODEL.DEVICE == "cuda"
    with pytest.raises(AttributeError):
        cfg.MODEL.NOT_A_KEY
    ref_dir = Path("/root/reference/configs")
    if ref_dir.exists():                                      # build container only: the shipped files load unchanged
        for path in sorted(ref_dir.glob("*.yaml")):
            node = ov.get_config(path)
            assert node.MODEL.VISION_EMBEDDING.D_FEATURE == 1024 and "ARCHITECTURE" in node.MODEL


def test_instance_list_zero_pads_ragged_inputs():
    a = ov.Instance(region_features=torch.ones(3, 8), region_boxes=torch.ones(3, 4), image_id=1)
    b = ov.Instance(region_features=torch.ones(5, 8), region_boxes=torch.ones(5, 4), image_id=2)
    items = ov.InstanceList([a, b])
    assert items.region_features.shape == (2, 5, 8) and float(items.region_features[0, 3:].abs().sum()) == 0
    assert items.batch_size == 2 and items.image_id == [1, 2] and items.missing is None
    assert items.to("cpu").region_boxes.shape == (2, 5, 4)
    assert ov.InstanceList().batch_size == 0


def test_synthetic_inputs_are_deterministic_and_bf16_exact():
    a = synthetic.synth_tensor("decoder.fc.weight", (50, 64), seed=3)
    b = synthetic.synth_tensor("decoder.fc.weight", (50, 64), seed=3)
    assert torch.equal(a, b) and torch.equal(a, a.to(torch.bfloat16).float())
    assert not torch.equal(a, synthetic.synth_tensor("decoder.fc.weight", (50, 64), seed=4))
    f = synthetic.synth_features(4, 10, 16, seed=1, ragged=True)
    assert float(f[0].abs().sum(-1).min()) > 0 and (f.sum(-1) == 0).any()
    vocab = synthetic.SyntheticVocab(20, 5)
    assert vocab.decode_caption(torch.tensor([[4, 5, 2, 0, 0]])) == ["w4 w5"]


def _header_functions():
    text = (REPO / "include" / "openviic_cap.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(cap_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol(cap_lib):
    declared = _header_functions()
    assert declared == set(cabi.SIGNATURES), declared ^ set(cabi.SIGNATURES)
    for name in declared:
        assert hasattr(cap_lib, name)
    assert cap_lib.cap_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", str(cabi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cap_[a-z0-9_]+)", out))
    assert declared <= exported


def test_argument_errors_surface_as_runtime_errors(cap_lib):
    """Validation happens before any CUDA work, so it is checkable without a GPU."""
    with pytest.raises(RuntimeError, match="cap_linear"):
        cabi.call("cap_linear", None, 8, None, None, None, 8, 0, 0, 4, 4, 8, None)
    with pytest.raises(RuntimeError, match="multiples of 8"):
        cabi.call("cap_linear", 16, 12, 16, None, 16, 8, 0, 0, 4, 4, 12, None)
    bad = cabi.ModelDesc(d_model=512, heads=8, d_k=32, d_v=32, d_ff=2048, d_feature=2048, enc_layers=3, dec_layers=3,
                         vocab=100, max_len=20, n_enc_levels=1)
    import ctypes as C
    handle = C.c_void_p()
    with pytest.raises(RuntimeError, match="d_k = d_v = 64"):
        cabi.call("cap_engine_create", C.byref(bad), C.byref(handle))


def test_no_cpu_fallback_and_no_oracle_in_the_product():
    cfg = ov.get_config("standard_transformer.yaml")
    cfg.MODEL.DEVICE = "cpu"
    vocab = synthetic.SyntheticVocab(50, 6)
    model = ov.build_model(cfg.MODEL, vocab)
    items = ov.InstanceList()
    items.set("grid_features", torch.zeros(1, 49, 2048))
    with pytest.raises(RuntimeError, match="CUDA"):
        model.beam_search(items, batch_size=1, beam_size=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ov.CaptionEngine(cfg.MODEL, vocab, model.state_dict(), "cpu")
    for path in (REPO / "openviic_b200").rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", path.read_text(), flags=re.M), path
    probe = "import sys; import openviic_b200; sys.exit(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    assert subprocess.run([sys.executable, "-c", probe], cwd=REPO).returncode == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", tmp_path / "libopenviic_cap.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cabi.load_library()


def test_reference_arm_prints_exactly_one_json_line():
    """`bench.py --impl reference`: stdout carries one JSON object with the contract's keys (the GPU arm prints the same
    shape); anything else the run prints goes to stderr."""
    proc = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-batch", "2"], capture_output=True, text=True, cwd=REPO, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "captions_per_sec_beam5_len20" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_gpu_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    proc = subprocess.run([sys.executable, str(REPO / "bench.py"), "--steps", "1"], capture_output=True, text=True, cwd=REPO,
                          timeout=600)
    assert proc.returncode != 0 and proc.stdout.strip() == ""
    assert "no CPU fallback" in proc.stderr
