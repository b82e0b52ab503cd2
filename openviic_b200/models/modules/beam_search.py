"""Beam search over any stateful model exposing ``step`` / ``apply_to_states``.

Same constructor and ``apply`` contract as the reference's models/modules/beam_search.py:5-118, but
select / EOS bookkeeping / history reordering run in the device state machine (``cap_beam_*``):
no full sort over beam*V, no host round trip per step.  The model's own states are reordered with
one ``index_select`` per state (the generic path; the engine avoids even that).
"""

from __future__ import annotations

import ctypes as C

import torch

from ... import cabi


class BeamSearch(object):
    # tests set this to a list to receive every step's (selected beams, selected words), each (b_s, beam), int64, and
    # (with return_probs) the final order of the beams as the last entry
    debug_trace = None

    def __init__(self, model, b_s: int, max_len: int, eos_idx: int, beam_size: int, device):
        self.model, self.b_s, self.max_len = model, b_s, max_len
        self.eos_idx, self.beam_size, self.device = eos_idx, beam_size, torch.device(device)

    def _expand_state(self, selected_rows: torch.Tensor):
        """Reorder a state whose dim 0 is (image, beam) rows (beam_search.py:19-34)."""
        def fn(s):
            return s.index_select(0, selected_rows)
        return fn

    def apply(self, out_size=1, return_probs=False, **kwargs):
        if self.device.type != "cuda":
            raise RuntimeError("BeamSearch runs on CUDA only (no CPU fallback)")
        b_s, beam, T = self.b_s, self.beam_size, self.max_len
        # return_probs (beam_search.py:68-81, 90, 103-118): the word log-probs of every step, masked like the
        # reference's (zero rows for finished beams), in the beam order OF THAT STEP (the reference never reorders the
        # earlier entries by later selections), gathered once by the final sort of the beams
        all_log_probs, seq_mask = [], torch.ones((b_s, beam, 1), device=self.device)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        handle = C.c_void_p()
        state = None
        try:
            selected_words = None
            for t in range(T):
                cur = 1 if t == 0 else beam
                word_logprob = self.model.step(t, selected_words, **kwargs).reshape(b_s, cur, -1).float()
                vocab = word_logprob.shape[-1]
                if state is None:
                    cabi.call("cap_beam_create", b_s, beam, T, vocab, self.eos_idx, C.byref(handle))
                    state = handle
                    cabi.call("cap_beam_reset", state, b_s, 0, stream)
                    lib = cabi.load_library()
                    tokens = self._view(lib.cap_beam_tokens(state), b_s * beam, torch.int32)
                    parents = self._view(lib.cap_beam_parents(state), b_s * beam, torch.int32)
                scores = word_logprob.expand(b_s, beam, vocab).contiguous() if cur == 1 else word_logprob.contiguous()
                if return_probs:
                    if t > 0:
                        seq_mask = seq_mask * (selected_words.view(b_s, cur) != self.eos_idx).float().unsqueeze(-1)
                        all_log_probs.append((word_logprob * seq_mask).unsqueeze(2))
                    else:
                        all_log_probs.append(scores.unsqueeze(2).clone())
                cabi.call("cap_beam_step", state, t, scores.data_ptr(), vocab, 1, stream)
                sel_beam = parents.view(b_s, beam).long()
                if return_probs:
                    seq_mask = torch.gather(seq_mask, 1, sel_beam.clone().unsqueeze(-1))
                base = torch.arange(b_s, device=self.device).view(-1, 1) * cur
                self.model.apply_to_states(self._expand_state((base + sel_beam).reshape(-1)))
                selected_words = tokens.long().view(-1, 1).clone()
                if BeamSearch.debug_trace is not None:
                    BeamSearch.debug_trace.append((sel_beam.clone(), selected_words.view(b_s, beam).clone()))
            ids = torch.empty((b_s, out_size, T), device=self.device, dtype=torch.int64)
            logp = torch.empty((b_s, out_size, T), device=self.device, dtype=torch.float32)
            cabi.call("cap_beam_finalize", state, out_size, ids.data_ptr(), logp.data_ptr(), stream)
            if return_probs:   # the same final order: beams by seq_logprob, descending, stable
                seq_lp = self._view(cabi.load_library().cap_beam_seq_logprob(state), b_s * beam, torch.float32).view(b_s, beam)
                order = torch.sort(seq_lp, dim=1, descending=True, stable=True).indices
                probs = torch.cat(all_log_probs, 2)
                if BeamSearch.debug_trace is not None:
                    BeamSearch.debug_trace.append(order.clone())
                probs = torch.gather(probs, 1, order.view(b_s, beam, 1, 1).expand(b_s, beam, T, probs.shape[-1]))
            torch.cuda.current_stream().synchronize()
        finally:
            if state is not None:
                cabi.call("cap_beam_destroy", state)
        if out_size == 1:
            ids, logp = ids.squeeze(1), logp.squeeze(1)
        if return_probs:
            return ids, logp, probs
        return ids, logp

    def _view(self, ptr: int, count: int, dtype: torch.dtype) -> torch.Tensor:
        """Zero-copy torch view of a device array owned by the beam handle."""
        itemsize = torch.empty((), dtype=dtype).element_size()

        class _Holder:
            pass

        holder = _Holder()
        holder.__cuda_array_interface__ = {
            "shape": (count,), "typestr": {torch.int32: "<i4", torch.float32: "<f4"}[dtype],
            "data": (int(ptr), False), "version": 2, "strides": (itemsize,)}
        return torch.as_tensor(holder, device=self.device)
