"""ctypes binding of include/openviic_cap.h (libopenviic_cap.so).

This is the only door between Python and the CUDA kernels.  There is deliberately no fallback:
if the shared library is missing or a call fails, a RuntimeError is raised.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libopenviic_cap.so"

CAP_BF16, CAP_F32 = 0, 1
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_LEAKY_RELU = 0, 1, 2, 3
ENC_PLAIN, ENC_MULTILEVEL, ENC_GEOMETRIC = 0, 1, 2
ATT_SDPA, ATT_GEOMETRY, ATT_MEMORY = 0, 1, 2
DEC_PLAIN, DEC_MESHED = 0, 1

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64


class AttentionArgs(C.Structure):
    """Mirror of ``cap_attention_args``."""

    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("out", _vp),
        ("q_bs", _i64), ("k_bs", _i64), ("v_bs", _i64), ("o_bs", _i64),
        ("ldq", _i), ("ldk", _i), ("ldv", _i), ("ldo", _i),
        ("mask", _vp), ("mask_bs", _i64), ("mask_qs", _i),
        ("geometry", _vp), ("mem_k", _vp), ("mem_v", _vp), ("n_mem", _i),
        ("B", _i), ("H", _i), ("nq", _i), ("nk", _i),
        ("scale", _f),
        ("sentinel", _vp), ("s_bs", _i64), ("lds", _i),
    ]


class ModelDesc(C.Structure):
    """Mirror of ``cap_model_desc``."""

    _fields_ = [(name, _i) for name in (
        "d_model", "heads", "d_k", "d_v", "d_ff", "d_feature", "enc_layers", "dec_layers",
        "encoder_kind", "enc_attention", "n_memory", "trig_geometry", "decoder_kind", "n_enc_levels",
        "aoa_enc", "aoa_dec_self", "aoa_dec_cross", "vocab", "max_len", "pad_idx", "bos_idx", "eos_idx")]


# name -> (restype, argtypes); must list every function declared in include/openviic_cap.h
SIGNATURES = {
    "cap_abi_version": (_i, []),
    "cap_last_error": (C.c_char_p, []),
    "cap_launch_count": (_i64, []),
    "cap_debug_gemm_trace": (_i, [_vp]),
    "cap_fault_records": (_i, [_vp, _i]),
    "cap_flight_records": (_i, [_vp, _i]),
    "cap_linear": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "cap_linear_simt": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "cap_add_layernorm": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _f, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "cap_feature_mask_cast": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp]),
    "cap_geometry_bias": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "cap_region_grid_mask": (_i, [_vp, _vp, _i, _i, _vp]),
    "cap_attention": (_i, [C.POINTER(AttentionArgs), _vp]),
    "cap_train_layernorm_fwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cap_train_layernorm_bwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cap_transpose_bf16": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _vp]),
    "cap_linear_splitk": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "cap_sum_partials": (_i, [_vp, _i, _i64, _vp, _vp]),
    "cap_train_relu_bwd": (_i, [_vp, _vp, _i64, _vp]),
    "cap_axpy_f32": (_i, [_vp, _vp, _i64, _vp]),
    "cap_attention_backward": (_i, [C.POINTER(AttentionArgs), _vp, _vp, _vp, _vp, _vp]),
    "cap_train_embed_fwd": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp]),
    "cap_train_embed_bwd": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    "cap_train_xent": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "cap_train_dropout": (_i, [_vp, _i, _i64, C.c_uint, _f, C.c_uint, C.c_uint, _vp]),
    "cap_train_adam": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _i, _vp]),
    "cap_decode_self_attention": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "cap_decode_cross_attention": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "cap_decode_cross_attention_levels": (_i, [_vp, _i, _vp, C.c_size_t, _vp, _vp, _i, C.c_size_t, _i, _i, _i, _i, _i, _f, _vp]),
    "cap_embed_tokens": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "cap_meshed_mix": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp]),
    "cap_aoa_gate": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cap_log_softmax": (_i, [_vp, _i, _vp, _i, _i, _i, _vp]),
    "cap_beam_create": (_i, [_i, _i, _i, _i, _i, C.POINTER(_vp)]),
    "cap_beam_destroy": (_i, [_vp]),
    "cap_beam_reset": (_i, [_vp, _i, _i, _vp]),
    "cap_beam_step": (_i, [_vp, _i, _vp, _i, _i, _vp]),
    "cap_vocab_logits_stats": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, C.POINTER(_i), _vp]),
    "cap_beam_step_stats": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp]),
    "cap_engine_decode_step": (_i, [_vp, _i, _vp]),
    "cap_beam_finalize": (_i, [_vp, _i, _vp, _vp, _vp]),
    "cap_beam_tokens": (_vp, [_vp]),
    "cap_beam_ancestry": (_vp, [_vp]),
    "cap_beam_seq_logprob": (_vp, [_vp]),
    "cap_beam_parents": (_vp, [_vp]),
    "cap_fused_weights_create": (_i, [_vp, _i, C.POINTER(_vp)]),
    "cap_fused_weights_destroy": (_i, [_vp]),
    "cap_fused_create": (_i, [_vp, C.POINTER(_vp)]),
    "cap_fused_destroy": (_i, [_vp]),
    "cap_debug_fused_trace": (_i, [_vp]),
    "cap_fused_set_full_logits": (_i, [_vp, _i]),
    "cap_fused_get_full_logits": (_i, [_vp]),
    "cap_fused_chain": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "cap_enc_chain_weights_create": (_i, [_vp, C.POINTER(_vp)]),
    "cap_enc_chain_weights_destroy": (_i, [_vp]),
    "cap_enc_chains_create": (_i, [_vp, C.POINTER(_vp)]),
    "cap_enc_chains_destroy": (_i, [_vp]),
    "cap_enc_chain": (_i, [_vp, _i, _i, _i, _vp]),
    "cap_linear_layernorm": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _f, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "cap_engine_create": (_i, [C.POINTER(ModelDesc), C.POINTER(_vp)]),
    "cap_engine_destroy": (_i, [_vp]),
    "cap_engine_create_shared": (_i, [_vp, C.POINTER(_vp)]),
    "cap_engine_load_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i]),
    "cap_engine_finalize": (_i, [_vp]),
    "cap_engine_reserve": (_i, [_vp, _i, _i, _i]),
    "cap_engine_encode": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp]),
    "cap_engine_decode_logits": (_i, [_vp, _i, _vp]),
    "cap_engine_beam_advance": (_i, [_vp, _i, _vp]),
    "cap_engine_begin_decode": (_i, [_vp, _vp]),
    "cap_engine_beam_search": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "cap_engine_caption_host": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "cap_engine_caption_host_async": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "cap_engine_caption_device_async": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "cap_engine_debug_chains": (_i, [_vp, _i, _vp]),
    "cap_host_collate_bf16": (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    "cap_host_collate_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    "cap_vocab_create": (_i, [_vp, _vp, _i64, _vp, _i64, C.POINTER(_vp)]),
    "cap_vocab_destroy": (_i, [_vp]),
    "cap_vocab_decode": (_i, [_vp, _vp, _i64, _i, _i, _vp, _i64, C.POINTER(_i64)]),
    "cap_cider_create": (_i, [_i, C.c_double, C.POINTER(_vp)]),
    "cap_cider_destroy": (_i, [_vp]),
    "cap_cider_set_corpus": (_i, [_vp, _vp, _vp, _vp, _i64, C.c_double, _vp, _i64]),
    "cap_cider_max_doc_freq": (_i64, [_vp]),
    "cap_cider_score": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, C.c_double, _vp, _i64, _vp, _i]),
    "cap_engine_encoder_output": (_vp, [_vp]),
    "cap_engine_encoder_mask": (_vp, [_vp]),
    "cap_engine_logits": (_vp, [_vp, C.POINTER(_i)]),
    "cap_engine_beam": (_vp, [_vp]),
}

# entry points whose int return value is an error code
_STATUS_FUNCS = {name for name, (res, _) in SIGNATURES.items()
                 if res is _i and name not in ("cap_abi_version", "cap_fault_records", "cap_flight_records",
                                              "cap_fused_get_full_logits")}

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """dlopen the CUDA library; fail loudly if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m openviic_b200.build` "
                "(or __graft_entry__.build()).  The caption path has no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    return load_library().cap_last_error().decode("utf-8", "replace")


def call(name: str, *args):
    """Invoke a status-returning entry point; raise RuntimeError carrying cap_last_error()."""
    lib = load_library()
    rc = getattr(lib, name)(*args)
    if name in _STATUS_FUNCS and rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {last_error()}")
    return rc


def launch_count() -> int:
    return int(load_library().cap_launch_count())


def fault_records():
    """Source lines (in their .cu file) of the device-side waits that timed out; readable after the CUDA context has
    faulted.  Empty: no bounded wait ever timed out, i.e. a fault had another cause."""
    buf = (C.c_uint * 64)()
    n = int(load_library().cap_fault_records(buf, 64))
    return [int(r) for r in buf[:min(n, 64)]]


FLIGHT_KINDS = ("chain", "chain_past_setup", "cross_producer", "cross_consumer", "cross_consumer_past_pdl_wait",
                "self_attention", "gemm", "gemm_past_setup", "attention", "beam_merge", "beam_select", "beam_other",
                "row_kernels")


def flight_records():
    """{kind: (CTAs entered, CTAs left)} from the flight recorder (OPENVIIC_FLIGHT=1), {} when it is off."""
    buf = (C.c_uint * 32)()
    n = int(load_library().cap_flight_records(buf, 16))
    return {name: (int(buf[2 * i]), int(buf[2 * i + 1])) for i, name in enumerate(FLIGHT_KINDS[:n])}
