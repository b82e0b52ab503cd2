"""Python handle of the whole-path engine (``cap_engine_*`` in include/openviic_cap.h)."""

from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import cabi

_ENCODERS = {"Encoder": cabi.ENC_PLAIN, "MultilevelEncoder": cabi.ENC_MULTILEVEL, "GeometricEncoder": cabi.ENC_GEOMETRIC}
_ATTENTIONS = {"ScaledDotProductAttention": cabi.ATT_SDPA,
               "AugmentedGeometryScaledDotProductAttention": cabi.ATT_GEOMETRY,
               "AugmentedMemoryScaledDotProductAttention": cabi.ATT_MEMORY}
_DECODERS = {"Decoder": cabi.DEC_PLAIN, "MeshedDecoder": cabi.DEC_MESHED}


def model_desc(model_cfg, vocab) -> cabi.ModelDesc:
    """Translate a MODEL config node into ``cap_model_desc``; raises ValueError for configs the
    engine does not cover (callers then use the generic module path)."""
    enc, dec = model_cfg.ENCODER, model_cfg.DECODER
    att = enc.SELF_ATTENTION
    d_self, d_cross = dec.ATTENTION.SELF_ATTENTION, dec.ATTENTION.ENC_ATTENTION
    if model_cfg.VISION_EMBEDDING.ARCHITECTURE != "FeatureEmbedding":
        raise ValueError("engine covers FeatureEmbedding only")
    if dec.TEXT_EMBEDDING.ARCHITECTURE != "UsualEmbedding" or dec.TEXT_EMBEDDING.get("WORD_EMBEDDING") is not None:
        raise ValueError("engine covers UsualEmbedding without pretrained vectors only")
    for name, table in ((enc.ARCHITECTURE, _ENCODERS), (att.ARCHITECTURE, _ATTENTIONS), (dec.ARCHITECTURE, _DECODERS)):
        if name not in table:
            raise ValueError(f"engine does not cover {name}")
    for a in (d_self, d_cross):
        if a.ARCHITECTURE != "ScaledDotProductAttention":
            raise ValueError("engine covers scaled dot-product attention in the decoder only")
        if (a.HEAD, a.D_KEY, a.D_VALUE, a.D_MODEL) != (att.HEAD, att.D_KEY, att.D_VALUE, att.D_MODEL):
            raise ValueError("engine needs the same head geometry in encoder and decoder")
    if d_cross.D_FF != att.D_FF or enc.D_MODEL != dec.D_MODEL:
        raise ValueError("engine needs the same D_FF / D_MODEL in encoder and decoder")
    meshed = dec.ARCHITECTURE == "MeshedDecoder"
    return cabi.ModelDesc(
        d_model=enc.D_MODEL, heads=att.HEAD, d_k=att.D_KEY, d_v=att.D_VALUE, d_ff=att.D_FF,
        d_feature=model_cfg.VISION_EMBEDDING.D_FEATURE, enc_layers=enc.LAYERS, dec_layers=dec.LAYERS,
        encoder_kind=_ENCODERS[enc.ARCHITECTURE], enc_attention=_ATTENTIONS[att.ARCHITECTURE],
        n_memory=int(att.get("MEMORY", 0) or 0) if att.ARCHITECTURE.startswith("AugmentedMemory") else 0,
        trig_geometry=int(bool(enc.get("TRIGNOMETRIC_EMBEDDING", False))),
        decoder_kind=_DECODERS[dec.ARCHITECTURE],
        n_enc_levels=int(dec.ATTENTION.N_ENCODER_LAYERS) if meshed else 1,
        aoa_enc=int(bool(att.USE_AOA)), aoa_dec_self=int(bool(d_self.USE_AOA)), aoa_dec_cross=int(bool(d_cross.USE_AOA)),
        vocab=len(vocab), max_len=vocab.max_caption_length, pad_idx=vocab.padding_idx, bos_idx=vocab.bos_idx,
        eos_idx=vocab.eos_idx)


def _feat_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return cabi.CAP_F32
    if t.dtype == torch.bfloat16:
        return cabi.CAP_BF16
    raise TypeError(f"features must be float32 or bfloat16, got {t.dtype}")


def _check_out(out, b: int, out_size: int, max_len: int, where: str, device=None):
    """(ids int64, log-probs float32), both contiguous (b, out_size, max_len): the C side writes exactly that many."""
    ids, logp = out
    for t, dt, name in ((ids, torch.int64, "ids"), (logp, torch.float32, "log-probs")):
        if t.dtype != dt or not t.is_contiguous() or t.numel() != b * out_size * max_len:
            raise ValueError(f"{where}: output {name} must be a contiguous {dt} tensor of {b}x{out_size}x{max_len} elements")
        if device is None and t.is_cuda:
            raise ValueError(f"{where}: output {name} must live in host memory")
        if device is not None and t.device != device:
            raise ValueError(f"{where}: output {name} must live on {device}")


class CaptionEngine:
    """One engine per GPU / rank: weights as bf16, workspaces, KV caches, beam state, CUDA graph."""

    def __init__(self, model_cfg, vocab, state_dict: Optional[Dict[str, torch.Tensor]], device="cuda",
                 share_weights_with: Optional["CaptionEngine"] = None):
        """``share_weights_with``: another engine of the same model on the same device whose device weights this one
        uses (``state_dict`` is then ignored): no upload, and concurrent engines read one L2-resident weight set."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CaptionEngine needs a CUDA device (there is no CPU fallback)")
        if self.device.index is None:   # "cuda" -> the current device, so that tensor.device comparisons are exact
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.desc = model_desc(model_cfg, vocab)
        self.max_len = vocab.max_caption_length
        self._h = C.c_void_p()
        self.reserved: Optional[Tuple[int, int, int]] = None
        if share_weights_with is not None:
            if share_weights_with.device != self.device:
                raise ValueError("share_weights_with: the parent engine lives on another device")
            with torch.cuda.device(self.device):
                cabi.call("cap_engine_create_shared", share_weights_with._h, C.byref(self._h))
            return
        with torch.cuda.device(self.device):
            cabi.call("cap_engine_create", C.byref(self.desc), C.byref(self._h))
            for name, tensor in state_dict.items():
                if not torch.is_tensor(tensor) or not tensor.dtype.is_floating_point or tensor.numel() == 0:
                    continue  # integer / empty decode-state buffers are not weights
                host = tensor.detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * host.dim())(*host.shape)
                cabi.call("cap_engine_load_weight", self._h, name.encode(), host.data_ptr(), shape, host.dim())
            cabi.call("cap_engine_finalize", self._h)

    def clone(self) -> "CaptionEngine":
        """A further engine over this engine's device weights, reserved like this one: for captioning independent
        batches concurrently on several streams."""
        other = CaptionEngine.__new__(CaptionEngine)
        other.device, other.desc, other.max_len = self.device, self.desc, self.max_len
        other._h, other.reserved = C.c_void_p(), None
        with torch.cuda.device(self.device):
            cabi.call("cap_engine_create_shared", self._h, C.byref(other._h))
        if self.reserved is not None:
            other.reserve(*self.reserved)
        return other

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            cabi.call("cap_engine_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reserve(self, max_batch: int, n_tokens: int, beam: int):
        with torch.cuda.device(self.device):
            cabi.call("cap_engine_reserve", self._h, max_batch, n_tokens, beam)
        self.reserved = (max_batch, n_tokens, beam)

    def fits(self, batch: int, n_tokens: int, beam: int) -> bool:
        return self.reserved is not None and batch <= self.reserved[0] and n_tokens <= self.reserved[1] \
            and beam == self.reserved[2]

    # -- the path ---------------------------------------------------------------------------
    def encode(self, feats: torch.Tensor, boxes: Optional[torch.Tensor] = None):
        """feats (B,n,D_FEATURE) fp32/bf16 on the device [+ boxes (B,n,4) fp32]."""
        feats, bx = self._device_inputs(feats, boxes)
        b, n, _ = feats.shape
        self._keep = (feats, bx)
        with torch.cuda.device(self.device):
            cabi.call("cap_engine_encode", self._h, feats.data_ptr(), _feat_dtype(feats),
                      None if bx is None else bx.data_ptr(), b, n, self._stream())
        self.batch, self.n = b, n

    def _device_inputs(self, feats: torch.Tensor, boxes: Optional[torch.Tensor]):
        """Features as contiguous fp32 / bf16 and boxes as contiguous fp32 (B,n,4) on this engine's device."""
        if feats.device != self.device:
            raise ValueError(f"features live on {feats.device}, the engine on {self.device}")
        if feats.dim() != 3 or feats.shape[2] != self.desc.d_feature:
            raise ValueError(f"features must be (B, n, {self.desc.d_feature}), got {tuple(feats.shape)}")
        if feats.dtype not in (torch.float32, torch.bfloat16):
            feats = feats.float()
        feats = feats.contiguous()
        bx = None
        if boxes is not None:
            if boxes.device != self.device or tuple(boxes.shape) != (feats.shape[0], feats.shape[1], 4):
                raise ValueError("boxes must be (B, n, 4) on the engine's device")
            bx = boxes.float().contiguous()
        return feats, bx

    def beam_search(self, out_size: int = 1, use_graph: bool = True):
        ids = torch.empty((self.batch, out_size, self.max_len), device=self.device, dtype=torch.int64)
        logp = torch.empty((self.batch, out_size, self.max_len), device=self.device, dtype=torch.float32)
        if use_graph:  # graph replay needs stable output addresses: decode into engine-owned buffers
            key = (self.batch, self.n, out_size)
            if getattr(self, "_graph_out", {}).get("key") != key:
                self._graph_out = {"key": key, "ids": ids, "logp": logp}
            g = self._graph_out
            with torch.cuda.device(self.device):
                cabi.call("cap_engine_beam_search", self._h, out_size, g["ids"].data_ptr(), g["logp"].data_ptr(), 1,
                          self._stream())
            return g["ids"].clone(), g["logp"].clone()
        with torch.cuda.device(self.device):
            cabi.call("cap_engine_beam_search", self._h, out_size, ids.data_ptr(), logp.data_ptr(), 0, self._stream())
        return ids, logp

    def caption_host(self, feats_host: torch.Tensor, boxes_host: Optional[torch.Tensor] = None, out_size: int = 1,
                     use_graph: bool = True, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                     sync: bool = True):
        """End to end from HOST tensors (pin them for full PCIe speed) to HOST ids / log-probs.
        ``sync=False`` returns right after enqueueing (synchronise the current stream before reading)."""
        if feats_host.is_cuda or feats_host.dim() != 3 or feats_host.shape[2] != self.desc.d_feature:
            raise ValueError(f"caption_host: features must be a host tensor (B, n, {self.desc.d_feature})")
        if feats_host.dtype not in (torch.float32, torch.bfloat16):   # any other dtype would be reinterpreted bit-wise
            feats_host = feats_host.float()
        feats_host = feats_host.contiguous()
        b, n, _ = feats_host.shape
        if boxes_host is not None:
            if boxes_host.is_cuda or tuple(boxes_host.shape) != (b, n, 4):
                raise ValueError("caption_host: boxes must be a host tensor (B, n, 4)")
            boxes_host = boxes_host.float().contiguous()
        if out is None:
            out = (torch.empty((b, out_size, self.max_len), dtype=torch.int64).pin_memory(),
                   torch.empty((b, out_size, self.max_len), dtype=torch.float32).pin_memory())
        _check_out(out, b, out_size, self.max_len, "caption_host")
        ids, logp = out
        self._keep_host = (feats_host, boxes_host, out)   # the async copies read / write these after the call returns
        with torch.cuda.device(self.device):
            cabi.call("cap_engine_caption_host" if sync else "cap_engine_caption_host_async", self._h, feats_host.data_ptr(),
                      _feat_dtype(feats_host), None if boxes_host is None else boxes_host.data_ptr(), b, n, out_size,
                      ids.data_ptr(), logp.data_ptr(), 1 if use_graph else 0, self._stream())
        self.batch, self.n = b, n
        return ids, logp

    def caption_device(self, feats: torch.Tensor, boxes: Optional[torch.Tensor] = None, out_size: int = 1,
                       use_graph: bool = True, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """Device tensors in, device tensors out, one C call (encode + beam search), nothing synchronises."""
        feats, bx = self._device_inputs(feats, boxes)
        b, n, _ = feats.shape
        if out is None:
            out = (torch.empty((b, out_size, self.max_len), device=self.device, dtype=torch.int64),
                   torch.empty((b, out_size, self.max_len), device=self.device, dtype=torch.float32))
        _check_out(out, b, out_size, self.max_len, "caption_device", self.device)
        self._keep = (feats, bx, out)
        with torch.cuda.device(self.device):
            cabi.call("cap_engine_caption_device_async", self._h, feats.data_ptr(), _feat_dtype(feats),
                      None if bx is None else bx.data_ptr(), b, n, out_size, out[0].data_ptr(), out[1].data_ptr(),
                      1 if use_graph else 0, self._stream())
        self.batch, self.n = b, n
        return out

    # -- step-wise / debug views ------------------------------------------------------------------
    def begin_decode(self):
        cabi.call("cap_engine_begin_decode", self._h, self._stream())

    def decode_logits(self, t: int) -> torch.Tensor:
        """Run the decoder stack for step t; returns a copy of the (R, V) fp32 logits."""
        cabi.call("cap_engine_decode_logits", self._h, t, self._stream())
        ld = C.c_int()
        ptr = cabi.load_library().cap_engine_logits(self._h, C.byref(ld))
        rows = self.batch * self.reserved[2]
        return _device_view(ptr, (rows, ld.value), torch.float32, self.device)[:, :self.desc.vocab].clone()

    def decode_step(self, t: int):
        """Production step t: decoder stack (one fused kernel when the model is covered) + vocabulary statistics
        + beam update."""
        cabi.call("cap_engine_decode_step", self._h, t, self._stream())

    def logits(self) -> torch.Tensor:
        """Copy of the (R, V) fp32 logits the last decode step left behind."""
        ld = C.c_int()
        ptr = cabi.load_library().cap_engine_logits(self._h, C.byref(ld))
        rows = self.batch * self.reserved[2]
        return _device_view(ptr, (rows, ld.value), torch.float32, self.device)[:, :self.desc.vocab].clone()

    def beam_advance(self, t: int):
        cabi.call("cap_engine_beam_advance", self._h, t, self._stream())

    def beam_tokens(self) -> torch.Tensor:
        lib = cabi.load_library()
        rows = self.batch * self.reserved[2]
        return _device_view(lib.cap_beam_tokens(lib.cap_engine_beam(self._h)), (rows,), torch.int32, self.device).clone()

    def beam_parents(self) -> torch.Tensor:
        lib = cabi.load_library()
        rows = self.batch * self.reserved[2]
        return _device_view(lib.cap_beam_parents(lib.cap_engine_beam(self._h)), (rows,), torch.int32, self.device).clone()

    def finalize(self, out_size: int = 1):
        lib = cabi.load_library()
        ids = torch.empty((self.batch, out_size, self.max_len), device=self.device, dtype=torch.int64)
        logp = torch.empty((self.batch, out_size, self.max_len), device=self.device, dtype=torch.float32)
        cabi.call("cap_beam_finalize", lib.cap_engine_beam(self._h), out_size, ids.data_ptr(), logp.data_ptr(),
                  self._stream())
        return ids, logp

    def encoder_output(self) -> torch.Tensor:
        """Copy of the encoder output: (B, n, d) bf16, or (B, L, n, d) for the multi-level encoder."""
        lib = cabi.load_library()
        d, layers = self.desc.d_model, self.desc.enc_layers
        cap_rows = self.reserved[0] * self.reserved[1]
        full = _device_view(lib.cap_engine_encoder_output(self._h), (layers, cap_rows, d), torch.bfloat16, self.device)
        used = full[:, : self.batch * self.n].reshape(layers, self.batch, self.n, d)
        if self.desc.encoder_kind == cabi.ENC_MULTILEVEL:
            return used.permute(1, 0, 2, 3).contiguous()
        return used[-1].clone()

    def encoder_mask(self) -> torch.Tensor:
        lib = cabi.load_library()
        return _device_view(lib.cap_engine_encoder_mask(self._h), (self.batch, self.n), torch.uint8, self.device).bool()


_TYPESTR = {torch.int32: "<i4", torch.float32: "<f4", torch.uint8: "|u1", torch.bfloat16: "<i2", torch.int64: "<i8"}


def _device_view(ptr: int, shape, dtype: torch.dtype, device) -> torch.Tensor:
    """Zero-copy torch view over device memory owned by the library."""
    class _Holder:
        pass

    holder = _Holder()
    holder.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": _TYPESTR[dtype], "data": (int(ptr), False),
                                       "version": 2, "strides": None}
    t = torch.as_tensor(holder, device=device)
    return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t
