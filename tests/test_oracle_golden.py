"""The oracle against the committed golden vectors (generated from HF transformers' Whisper, the one
independent implementation importable offline -- tests/golden/make_golden.py) and its own invariants."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_weights

G = Path(__file__).parent / "golden"


@pytest.mark.parametrize("n_mels", [80, 128])
def test_logmel_matches_hf_feature_extractor(n_mels):
    z = np.load(G / f"hf_logmel_{n_mels}.npz")
    a = speech_shaped_audio(12.0, 1001)
    ap = np.zeros(480000, np.float32)
    ap[: len(a)] = a
    mel = wo.log_mel_spectrogram(ap, n_mels).numpy()  # HF semantics: pad raw audio to 30 s, then STFT
    assert np.abs(mel[:, z["frames"]] - z["mel"]).max() <= 1e-5


def test_mel_filters_match_product_copy():
    from whisperjav_b200.model import slaney_mel_filters
    for n in (80, 128):
        assert np.abs(slaney_mel_filters(n) - wo.mel_filters(n)).max() <= 1e-7


def test_tiny_encoder_decoder_match_hf():
    z = np.load(G / "hf_tiny_greedy.npz")
    d = DIMS["tiny"]
    w = synth_weights(d, seed=7)
    a = speech_shaped_audio(12.0, 1001)
    ap = np.zeros(480000, np.float32)
    ap[: len(a)] = a
    mel = wo.log_mel_spectrogram(ap, 80)[None]
    enc = wo.encoder_forward(w, d, mel, sim_fp16=False)
    assert np.abs(enc[0, ::50, ::16].numpy() - z["enc"]).max() <= 2e-4
    # unfiltered greedy continuation, fp32, full-recompute each step in HF vs kv-cached oracle
    ids = z["ids"].tolist()
    st = wo.DecoderState()
    toks = torch.tensor([ids[:4]])
    out = list(ids[:4])
    for step in range(12):
        lg = wo.decoder_forward(w, d, toks, enc, st, sim_fp16=False)[0, -1]
        assert np.abs(lg[::97].numpy() - z["logits"][step]).max() <= 2e-3
        nxt = int(lg.argmax())
        out.append(nxt)
        toks = torch.tensor([[nxt]])
    assert out == ids


def test_sim_fp16_is_close_to_fp32():
    d = DIMS["tiny"]
    w = synth_weights(d, seed=7)
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(speech_shaped_audio(4.0, 5), 80, padding=wo.N_SAMPLES)[:, :400], 3000)[None]
    a = wo.encoder_forward(w, d, mel, sim_fp16=True)
    b = wo.encoder_forward(w, d, mel, sim_fp16=False)
    assert ((a - b).norm() / b.norm()).item() < 2e-2  # fp16 rounding noise through 4 peaky-softmax layers


def test_decode_rules_and_nondegenerate_trajectories():
    d = DIMS["tiny"]
    w = wo.prepare_weights(synth_weights(d, seed=7), True)
    clips = [speech_shaped_audio(s, 1000 + i) for i, s in enumerate([6.0, 3.0])]
    mel = torch.stack([wo.pad_or_trim(wo.log_mel_spectrogram(c, 80, padding=wo.N_SAMPLES)[:, : len(c) // 160], 3000) for c in clips])
    res = wo.decode(w, d, mel, wo.DecodingOptions(language="ja", max_initial_timestamp=0.0, sample_len=48), True)
    tok = wo.SpecialTokens(d.n_vocab, language="ja")
    supp = set(wo.get_suppress_tokens(tok, wo.DecodingOptions()))
    for r in res:
        t = r.tokens
        assert t[0] >= tok.timestamp_begin                      # first sampled token is a timestamp
        assert not (set(t) & supp) and tok.no_timestamps not in t
        ts = [x for x in t if x >= tok.timestamp_begin]
        assert ts == sorted(ts)                                  # timestamps never decrease
        assert len(set(t)) >= min(len(t), 8) // 2                # not a constant trajectory
        assert np.isfinite(r.avg_logprob) and 0.0 <= r.no_speech_prob <= 1.0
    assert res[0].tokens != res[1].tokens                        # depends on the audio


def test_transcribe_seek_loop_shapes():
    d = DIMS["tiny"]
    w = wo.prepare_weights(synth_weights(d, seed=7), True)
    out = wo.transcribe(w, d, speech_shaped_audio(7.0, 3), language="ja", temperature=0.0, condition_on_previous_text=False,
                        max_initial_timestamp=0.0, sample_len=24)
    assert out["language"] == "ja" and isinstance(out["segments"], list)
    for s in out["segments"]:
        assert s["end"] >= s["start"] >= 0.0 and {"id", "seek", "tokens", "avg_logprob", "no_speech_prob", "compression_ratio"} <= set(s)
    # empty audio -> no windows
    assert wo.transcribe(w, d, np.zeros(0, np.float32), language="ja")["segments"] == []
