"""Base captioning model: ``step`` + ``beam_search`` (reference: models/base_transformer.py:8-53).

``beam_search`` keeps the reference's signature and return values.  When the configuration is one
the whole-path engine covers (every shipped config on BASELINE.json's list) it runs there:
encoder once, cross K/V projected once per image, ancestry-indirected self-attention cache, fused
log-softmax/top-k, the 20-step loop replayed from a CUDA graph.  Otherwise it falls back to the
generic *CUDA* path (``BeamSearch`` over ``step`` and the registered modules) -- never to the CPU.
"""

from __future__ import annotations

import torch

from ..engine import CaptionEngine
from ..utils.instance import InstanceList
from .modules.beam_search import BeamSearch
from .modules.containers import Module


class BaseTransformer(Module):
    def __init__(self, vocab):
        super().__init__()
        self.vocab = vocab
        self.max_len = vocab.max_caption_length
        self.eos_idx = vocab.eos_idx
        self.register_state("encoder_features", None)
        self.register_state("encoder_padding_mask", None)
        self.model_config = None
        self._engine = None
        self._engine_key = None
        self.use_cuda_graph = True

    # ---- to be provided by the architectures ----
    def encoder_forward(self, input_features: InstanceList):
        raise NotImplementedError

    def engine_inputs(self, input_features: InstanceList):
        """(features, boxes-or-None) the engine consumes for this architecture."""
        raise NotImplementedError

    def forward(self, input_features: InstanceList):
        """Teacher-forced log-probs (B,T,V) (standard_stransformer.py:21-31)."""
        encoder_features, encoder_padding_mask = self.encoder_forward(input_features)
        return self.decoder(caption_tokens=input_features.caption_tokens, encoder_features=encoder_features,
                            encoder_attention_mask=encoder_padding_mask)

    def step(self, t, prev_output):
        bs = self.encoder_features.shape[0]
        if t == 0:
            it = torch.full((bs, 1), self.vocab.bos_idx, dtype=torch.long, device=self.encoder_features.device)
        else:
            it = prev_output
        return self.decoder(caption_tokens=it, encoder_features=self.encoder_features,
                            encoder_attention_mask=self.encoder_padding_mask)

    # ---- engine plumbing ----
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, batch_size: int, n_tokens: int, beam_size: int) -> CaptionEngine:
        """The (lazily built, cached) whole-path engine for this model on its device."""
        key = self._weights_version()
        stale = self._engine is None or self._engine_key != key
        if stale or not self._engine.fits(batch_size, n_tokens, beam_size):
            if stale:
                if self._engine is not None:
                    self._engine.close()
                self._engine = CaptionEngine(self.model_config, self.vocab, self.state_dict(),
                                             next(self.parameters()).device)
                self._engine_key = key
            else:
                # same weights, larger shapes: rebuild (reservations are fixed-size) at the element-wise maximum of the
                # old and the new request, so that batches of ragged sizes grow the reservation monotonically instead
                # of rebuilding the engine every time n or the batch changes
                old = self._engine.reserved
                if old is not None and old[2] == beam_size:
                    batch_size, n_tokens = max(batch_size, old[0]), max(n_tokens, old[1])
                self._engine.close()
                self._engine = CaptionEngine(self.model_config, self.vocab, self.state_dict(),
                                             next(self.parameters()).device)
            self._engine.reserve(batch_size, n_tokens, beam_size)
        return self._engine

    def engine_supported(self) -> bool:
        from ..engine import model_desc
        try:
            model_desc(self.model_config, self.vocab)
            return True
        except (ValueError, AttributeError):
            return False

    def beam_search(self, input_features: InstanceList, batch_size: int, beam_size: int, out_size=1,
                    return_probs=False, **kwargs):
        device = next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("openviic_b200 models decode on a CUDA device only (no CPU fallback)")
        # `disable_engine` (set by the training loop while the weights change every iteration) keeps decoding on the
        # module-level CUDA path, whose bf16 weight copies follow the parameters' versions without an engine rebuild
        if not return_probs and not kwargs and not getattr(self, "disable_engine", False) and self.engine_supported():
            feats, boxes = self.engine_inputs(input_features)
            eng = self.engine(batch_size, feats.shape[1], beam_size)
            eng.encode(feats.to(device), None if boxes is None else boxes.to(device))
            ids, logp = eng.beam_search(out_size, use_graph=self.use_cuda_graph)
            if out_size == 1:
                ids, logp = ids.squeeze(1), logp.squeeze(1)
            return ids, logp
        beam_search = BeamSearch(model=self, max_len=self.max_len, eos_idx=self.eos_idx, beam_size=beam_size,
                                 b_s=batch_size, device=device)
        with self.statefulness(batch_size):
            self.encoder_features, self.encoder_padding_mask = self.encoder_forward(input_features)
            return beam_search.apply(out_size, return_probs, **kwargs)

    def generic_beam_search(self, input_features: InstanceList, batch_size: int, beam_size: int, out_size=1):
        """The module-level path (``step`` + ``BeamSearch``), kept callable for parity tests."""
        device = next(self.parameters()).device
        beam_search = BeamSearch(model=self, max_len=self.max_len, eos_idx=self.eos_idx, beam_size=beam_size,
                                 b_s=batch_size, device=device)
        with self.statefulness(batch_size):
            self.encoder_features, self.encoder_padding_mask = self.encoder_forward(input_features)
            return beam_search.apply(out_size, False)
