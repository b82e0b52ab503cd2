"""Beam-search state machine: bit-exact selection against the oracle's sort-based restatement."""

import ctypes as C

import pytest
import torch

from openviic_b200 import cabi
from openviic_b200.models.modules.beam_search import BeamSearch
from oracle import caption_oracle as oracle
from kat import ScriptedModel, scripted_kat

pytestmark = pytest.mark.gpu


def _scripted_scores(T, B, beam, V, eos, seed, quantum):
    """(T, R, V) fp32 'log-probs' full of exact ties (multiples of `quantum`) with <eos> often on top."""
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(T, B * beam, V, generator=g) * 2 - 6
    s = torch.round(s / quantum) * quantum
    boost = torch.rand(T, B * beam, generator=g) < 0.25
    s[..., eos] = torch.where(boost, torch.full_like(s[..., eos], -0.5), s[..., eos])
    return s


def _oracle_run(scores, B, beam, eos, out_size, stable_ties=True):
    T = scores.shape[0]
    trace = []

    def step(t, prev):
        return (scores[t][::beam] if t == 0 else scores[t]).unsqueeze(1)

    ids, lp = oracle.beam_search(step, lambda fn: None, B, beam, T, eos, out_size, trace, stable_ties=stable_ties)
    return ids, lp, trace


def _cuda_run(scores, B, beam, eos, out_size, device, is_logprob=1):
    T, R, V = scores.shape
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = C.c_void_p()
    cabi.call("cap_beam_create", B, beam, T, V, eos, C.byref(h))
    try:
        cabi.call("cap_beam_reset", h, B, 1, stream)
        dev_scores = scores.to(device).contiguous()
        for t in range(T):
            cabi.call("cap_beam_step", h, t, dev_scores[t].data_ptr(), V, is_logprob, stream)
        ids = torch.empty(B, out_size, T, dtype=torch.int64, device=device)
        lp = torch.empty(B, out_size, T, dtype=torch.float32, device=device)
        cabi.call("cap_beam_finalize", h, out_size, ids.data_ptr(), lp.data_ptr(), stream)
        torch.cuda.synchronize()
        lib = cabi.load_library()
        from openviic_b200.engine import _device_view
        anc = _device_view(lib.cap_beam_ancestry(h), (T, R), torch.int32, device).clone().cpu()
        seq = _device_view(lib.cap_beam_seq_logprob(h), (R,), torch.float32, device).clone().cpu()
    finally:
        cabi.call("cap_beam_destroy", h)
    return ids.cpu(), lp.cpu(), anc, seq


@pytest.mark.parametrize("B,beam,V,T,quantum", [(9, 5, 1000, 20, 0.25), (4, 3, 10201, 20, 0.5), (16, 5, 51, 12, 1.0),
                                                (3, 8, 300, 20, 0.25), (5, 1, 97, 9, 0.5)])
def test_beam_steps_bit_exact_with_ties_and_eos(device, B, beam, V, T, quantum):
    eos = 2
    scores = _scripted_scores(T, B, beam, V, eos, seed=B * 100 + V, quantum=quantum)
    ref_ids, ref_lp, trace = _oracle_run(scores, B, beam, eos, beam)
    ids, lp, anc, seq = _cuda_run(scores, B, beam, eos, beam, device)
    assert torch.equal(ids, ref_ids.view(B, beam, T))
    assert torch.equal(lp, ref_lp.view(B, beam, T))          # same single fp32 addition per candidate
    if V <= 1000:
        assert (ref_ids == eos).any() and (ref_ids == 0).any()   # the case really exercises finished beams
    # ancestry table == explicit replay of the reference's per-step state gathers
    R = B * beam
    expect = torch.arange(R).repeat(T, 1)
    row0 = (torch.arange(R) // beam) * beam
    for t, rec in enumerate(trace):
        parent_rows = row0 + rec["beam"].reshape(-1)
        expect[:t] = expect[:t, parent_rows]
        expect[t] = parent_rows
    assert torch.equal(anc.long(), expect)
    assert torch.equal(seq, trace[-1]["seq_logprob"].reshape(-1))


@pytest.mark.parametrize("B,beam,V,T", [(11, 5, 10201, 20), (6, 3, 1000, 20)])
def test_beam_steps_bit_exact_with_reference_sort_as_written(device, B, beam, V, T):
    """Tie-free scores: the reference's own (unstable) torch.sort call selects identical beams."""
    eos = 2
    g = torch.Generator().manual_seed(V + B)
    scores = torch.log_softmax(torch.randn(T, B * beam, V, generator=g) * 3, -1)
    scores[..., eos] += torch.where(torch.rand(T, B * beam, generator=g) < 0.2, 6.0, 0.0)
    ref_ids, ref_lp, _ = _oracle_run(scores, B, beam, eos, beam, stable_ties=False)
    ids, lp, _, _ = _cuda_run(scores, B, beam, eos, beam, device)
    assert torch.equal(ids, ref_ids.view(B, beam, T)) and torch.equal(lp, ref_lp.view(B, beam, T))
    assert (ref_ids == eos).any()


def test_fused_log_softmax_path_matches_oracle(device):
    B, beam, V, T, eos = 6, 5, 1000, 15, 2
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(T, B * beam, V, generator=g) * 3
    logits[..., eos] += 4.0
    ref_ids, ref_lp, _ = _oracle_run(torch.log_softmax(logits, -1), B, beam, eos, 1)
    ids, lp, _, _ = _cuda_run(logits, B, beam, eos, 1, device, is_logprob=0)
    assert torch.equal(ids.squeeze(1), ref_ids)
    assert (lp.squeeze(1) - ref_lp).abs().max().item() < 1e-4


def test_scripted_model_kat_through_beamsearch_class(device):
    """The survey's scripted-model known answer, driven through the reference-shaped BeamSearch API."""
    model = ScriptedModel(device)
    bs = BeamSearch(model, b_s=1, max_len=4, eos_idx=2, beam_size=3, device=device)
    ids, lp = bs.apply(out_size=3)
    exp_ids, exp_lp, exp_fed = scripted_kat()
    assert ids.cpu().tolist() == exp_ids
    assert torch.allclose(lp.cpu(), torch.tensor(exp_lp), atol=1e-6)
    assert model.fed == exp_fed


def test_beam_argument_errors(device):
    h = C.c_void_p()
    with pytest.raises(RuntimeError, match="beam must be in"):
        cabi.call("cap_beam_create", 4, 9, 20, 100, 2, C.byref(h))
    cabi.call("cap_beam_create", 4, 5, 20, 100, 2, C.byref(h))
    try:
        with pytest.raises(RuntimeError, match="outside"):
            cabi.call("cap_beam_reset", h, 5, 1, None)
    finally:
        cabi.call("cap_beam_destroy", h)


@pytest.mark.parametrize("B,beam,V", [(7, 5, 10201), (5, 3, 1000), (3, 8, 777), (4, 5, 97)])
def test_vocab_stats_epilogue_matches_full_row_pass(device, B, beam, V):
    """Vocabulary GEMM with chunk statistics + chunk merge == GEMM -> logits -> full row pass, including
    exact ties that straddle chunks (quantised weights) and rows that finish (<eos>)."""
    from openviic_b200 import ops
    from openviic_b200.engine import _device_view
    T, d, eos = 6, 512, 2
    R = B * beam
    g = torch.Generator().manual_seed(V)
    w = torch.round(torch.randn(V, d, generator=g) * 0.13 * 8) / 8         # coarse grid: many exactly equal logits
    w[eos] = 0.25
    w = w.to(torch.bfloat16).to(device)
    xs = [torch.round(torch.randn(R, d, generator=g) * 2) / 2 for _ in range(T)]
    xs = [x.to(torch.bfloat16).to(device) for x in xs]
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = cabi.load_library()
    ld = (V + 7) // 8 * 8
    chunks = ((V + 255) // 256) * 8
    logits = torch.empty(R, ld, device=device)
    part_ms = torch.empty(R, chunks, 2, device=device)
    results = []
    for fused in (False, True):
        h = C.c_void_p()
        cabi.call("cap_beam_create", B, beam, T, V, eos, C.byref(h))
        try:
            cabi.call("cap_beam_reset", h, B, 1, stream)
            for t in range(T):
                if fused:
                    n_chunks = C.c_int()
                    cabi.call("cap_vocab_logits_stats", xs[t].data_ptr(), d, w.data_ptr(), None, logits.data_ptr(), ld,
                              R, V, d, part_ms.data_ptr(), C.byref(n_chunks), stream)
                    assert n_chunks.value == chunks
                    cabi.call("cap_beam_step_stats", h, t, logits.data_ptr(), ld, part_ms.data_ptr(), chunks, stream)
                else:
                    full = ops.linear(xs[t], w, None, out_dtype=torch.float32).contiguous()
                    cabi.call("cap_beam_step", h, t, full.data_ptr(), full.stride(0), 0, stream)
            ids = torch.empty(B, beam, T, dtype=torch.int64, device=device)
            lp = torch.empty(B, beam, T, dtype=torch.float32, device=device)
            cabi.call("cap_beam_finalize", h, beam, ids.data_ptr(), lp.data_ptr(), stream)
            torch.cuda.synchronize()
            seq = _device_view(lib.cap_beam_seq_logprob(h), (R,), torch.float32, device).clone().cpu()
            results.append((ids.cpu(), lp.cpu(), seq))
        finally:
            cabi.call("cap_beam_destroy", h)
    (ids_a, lp_a, seq_a), (ids_b, lp_b, seq_b) = results
    assert torch.equal(ids_a, ids_b)
    assert (lp_a - lp_b).abs().max().item() < 1e-4 and (seq_a - seq_b).abs().max().item() < 1e-4
    # statistics against a direct fp32 statement of the last step
    ref = xs[-1].float() @ w.float().t()
    assert (logits[:, :V] - ref).abs().max().item() < 1e-3
    pm, ps = part_ms[..., 0].cpu(), part_ms[..., 1].cpu()
    mx = pm.max(1).values
    mine = mx + torch.log((ps * torch.exp(pm - mx[:, None])).sum(1))
    assert (mine - torch.logsumexp(ref, -1).cpu()).abs().max().item() < 1e-3
