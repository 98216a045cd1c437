"""Parity at the benchmarked configuration: whisper-large-v3 (32 x 1280, 128 mels, vocabulary 51866), the model bench.py times.
Same checker as the tiny tests (oracle/parity.py) on the large-v3 synthetic weights bench.py uses (synth_preset("large-v3")):

  (i)   log-mel (128 bins) <= 1e-3 abs vs the oracle;
  (ii)  encoder hidden states <= 1e-2 relative vs ``encoder_forward(sim_fp16=True)`` fed the same fp16 mel, also after blocks
        8 / 16 / 24 / 32 (residual-stream taps, ``wjb_encoder_set_tap``);
  (iii) cross-attention K/V of every layer vs the oracle's Linear on the same encoder output;
  (iv)  decode in both timestamp modes: every step's raw logits vs the oracle teacher-forced along the device's sequence
        (max over the vocabulary <= 40 fp16 quanta, rms <= 8), every device token the oracle's arg-max on that prefix or a counted
        near-tie (<= 8 quanta), the
        large-v3 special-token ids (timestamp_begin 50365 ...) exercised through the filters.

The oracle needs ~0.1 s per decoder token per window on the GPU box's host cores, so the horizon is SAMPLE_LEN tokens on
N_WIN windows (the batch-64 / BN=256 tile shapes of the bench are exercised separately by ``test_batch64_rows_agree``)."""
import json

import pytest
import torch

from oracle import parity as P
from oracle import whisper_oracle as wo
from whisperjav_b200 import model as M
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_preset, synth_weights

pytestmark = pytest.mark.gpu
N_WIN = 3
SAMPLE_LEN = 48


@pytest.fixture(scope="module")
def large():
    dims = DIMS["large-v3"]
    w = synth_weights(dims, **synth_preset("large-v3"))
    m = M.WhisperB200(dims, w, max_batch=64)
    return dims, w, m, wo.prepare_weights(w, True)


@pytest.fixture(scope="module")
def clips():
    return [speech_shaped_audio(s, 2000 + i) for i, s in enumerate([30.0, 17.3, 30.0][:N_WIN])]


@pytest.fixture(scope="module")
def encoded(large, clips):
    dims, w, m, pw = large
    mel = P.gpu_mel(m, clips)
    rep, xa = P.encoder_parity(m, w, dims, mel, tap_every=8, prepared=pw)
    return mel, rep, xa


def test_mel_128(large, clips):
    dims, w, m, pw = large
    mel_tm = P.gpu_mel(m, clips)
    got = mel_tm[:, 1:-1].permute(0, 2, 1).float().cpu()
    assert (got - P.oracle_mel_windows(clips, dims)).abs().max().item() <= 1e-3


def test_encoder_hidden_states(large, encoded, diag_dir):
    mel, rep, xa = encoded
    (diag_dir / "encoder_large_v3.json").write_text(json.dumps(rep, indent=1))
    assert [t["after_block"] for t in rep["taps"]] == [8, 16, 24, 32]
    assert rep["ok"], rep            # <= 1e-2 relative at every tap and at the output
    assert rep["per_row_rel_max"] <= 5e-2, rep


def test_cross_kv(large, encoded):
    dims, w, m, pw = large
    _, _, xa = encoded
    B, H, T = xa.shape[0], dims.n_text_head, dims.n_audio_ctx
    from whisperjav_b200 import _lib
    kv = torch.empty(m.lib.wjb_cross_kv_bytes(m._h, B) // 2, dtype=torch.float16, device=m.device)
    _lib.check(m.lib.wjb_cross_kv(m._h, _lib.ptr(xa), B, _lib.ptr(kv), _lib.stream_ptr()), "wjb_cross_kv")
    kv = kv.view(dims.n_text_layer, B, 2 * H, T, 64).float().cpu()
    xin = xa.float().cpu()
    r = wo.Rounder(True)
    for layer in (0, 13, 31):
        p = f"decoder.blocks.{layer}.cross_attn"
        k = wo._linear(xin, pw, p + ".key", r).view(B, T, H, 64).permute(0, 2, 1, 3)
        v = wo._linear(xin, pw, p + ".value", r).view(B, T, H, 64).permute(0, 2, 1, 3)
        for got, ref in ((kv[layer, :, :H], k), (kv[layer, :, H:], v)):
            # same fp16 inputs, fp32 accumulation in a different order: at most an ulp of the output
            assert (got - ref).abs().max().item() <= 2.0 ** -9 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("without_timestamps", [False, True])
def test_decode_logits_and_tokens(large, encoded, diag_dir, without_timestamps):
    dims, w, m, pw = large
    _, _, xa = encoded
    # tolerances: a 32-layer random-init decoder amplifies a perturbation of fp16-rounding size (3e-4 relative on the encoder
    # output) to ~20 quanta of logit difference by itself (scripts/synth_chaos.py: max 19, median 9 over 45 steps with this
    # preset; the 4-layer tiny model: 8 / 4), so the bounds are 2.5 x the tiny ones
    rep = P.decode_parity(m, w, dims, xa, prepared=pw, language="ja", without_timestamps=without_timestamps, max_initial_timestamp=0.0,
                          sample_len=SAMPLE_LEN, tie_quanta=8.0, logit_quanta=40.0, logit_rms_quanta=8.0, logprob_tol_per_step=0.06)
    (diag_dir / f"tokens_large_v3_wt{int(without_timestamps)}.json").write_text(json.dumps(rep, indent=1))
    assert rep["ok"], rep["failures"]
    assert rep["steps_checked"] >= N_WIN * 8
    assert rep["tie_breaks"] <= max(2, rep["steps_checked"] // 25), rep
    assert rep["identical_windows"] >= 1, rep
    if not without_timestamps:
        tsb = M.Tokens(dims.n_vocab, "ja").timestamp_begin
        assert tsb == 50365 and all(t[0] >= tsb for t in rep["tokens"])      # large-v3 offsets went through the filters


def test_batch64_rows_agree(large, clips):
    """The benchmarked shapes (batch 64: BN=256 encoder tiles, 64-row step GEMMs, 64 x 20 attention CTAs) against the small batch
    checked against the oracle above: replicating the windows to 64 rows must not change any row's output."""
    dims, w, m, pw = large
    mel = P.gpu_mel(m, clips)
    xa_small = m.encode(mel)
    mel64 = mel[[i % N_WIN for i in range(64)]].contiguous()
    xa64 = m.encode(mel64)
    for i in range(64):
        assert torch.equal(xa64[i], xa_small[i % N_WIN]), i               # tiling of the batch does not change a row
    kw = dict(language="ja", without_timestamps=True, sample_len=32)
    small = m.decode_features(xa_small, **kw)
    big = m.decode_features(xa64, **kw)
    for i in range(64):
        assert big[i].tokens == small[i % N_WIN].tokens, i
        assert abs(big[i].sum_logprob - small[i % N_WIN].sum_logprob) <= 1e-3
