"""The oracle against the committed golden vectors (generated from HF transformers' Whisper, the one
independent implementation importable offline -- tests/golden/make_golden.py) and its own invariants."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo
from whisperjav_b200.synth import DIMS, Dims, speech_shaped_audio, synth_preset, synth_weights

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
    w = synth_weights(d, **synth_preset("tiny"))
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


@pytest.mark.parametrize("tag,n_vocab", [("v2", 51865), ("v3", 51866)])
@pytest.mark.parametrize("mode", ["ts", "ts_mi50", "nots"])
def test_logit_filters_match_hf_processors(tag, n_vocab, mode):
    """SuppressBlank / SuppressTokens / ApplyTimestampRules of the oracle against HF's own logits processors driving an HF
    model (fixture: make_golden.py::filters_fixture): same greedy ids and the same number of masked logits at every step, with
    the 51865-token and the 51866-token (large-v3) special-token ids."""
    z = np.load(G / f"hf_filters_{tag}vocab.npz")
    ids, masked = z[mode + "_ids"].tolist(), z[mode + "_masked"].tolist()
    d0 = DIMS["tiny"]
    d = Dims(**{**d0.__dict__, "n_vocab": n_vocab})
    w = synth_weights(d, **synth_preset("tiny"))
    a = speech_shaped_audio(12.0, 1001)
    ap = np.zeros(480000, np.float32)
    ap[: len(a)] = a
    enc = wo.encoder_forward(w, d, wo.log_mel_spectrogram(ap, 80)[None], sim_fp16=False)
    opts = wo.DecodingOptions(language="ja", without_timestamps=(mode == "nots"), max_initial_timestamp=1.0 if mode == "ts_mi50" else None)
    tok = wo.SpecialTokens(n_vocab, language="ja")
    initial = wo.get_initial_tokens(tok, opts, d.n_text_ctx)
    assert list(initial) == ids[: len(initial)]
    suppress = wo.get_suppress_tokens(tok, opts)
    mi = 50 if mode == "ts_mi50" else None
    st = wo.DecoderState()
    toks = torch.tensor([ids[: len(initial)]])
    for step in range(len(masked)):
        inp = toks if step == 0 else toks[:, -1:]
        lg = wo.decoder_forward(w, d, inp, enc, st, sim_fp16=False)[:, -1]
        wo.apply_logit_filters(lg, toks, tok, opts, len(initial), suppress, mi)
        assert int(torch.isinf(lg[0]).sum()) == masked[step], (step, "mask size")
        assert int(lg[0].argmax()) == ids[len(initial) + step], (step, "token")
        toks = torch.cat([toks, torch.tensor([[ids[len(initial) + step]]])], dim=1)


def test_timing_oracle_matches_hf():
    """oracle/timing_oracle.py median filter and DTW against HF's `_median_filter` / `_dynamic_time_warping`."""
    from oracle import timing_oracle as to
    z = np.load(G / "hf_timing.npz")
    got = to.median_filter(torch.from_numpy(z["med_x"]), 7).numpy()
    assert np.array_equal(got, z["med_y"])
    for i in range(4):
        ti, fi = to.dtw(-z[f"dtw{i}_x"].astype(np.float64))
        assert np.array_equal(ti, z[f"dtw{i}_text"]) and np.array_equal(fi, z[f"dtw{i}_time"]), i


def test_sim_fp16_is_close_to_fp32():
    d = DIMS["tiny"]
    w = synth_weights(d, **synth_preset("tiny"))
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(speech_shaped_audio(4.0, 5), 80, padding=wo.N_SAMPLES)[:, :400], 3000)[None]
    a = wo.encoder_forward(w, d, mel, sim_fp16=True)
    b = wo.encoder_forward(w, d, mel, sim_fp16=False)
    assert ((a - b).norm() / b.norm()).item() < 3e-3  # fp16 rounding noise through 4 layers (soft encoder attention, see synth.py)


def test_decode_rules_and_nondegenerate_trajectories():
    d = DIMS["tiny"]
    w = wo.prepare_weights(synth_weights(d, **synth_preset("tiny")), True)
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
    w = wo.prepare_weights(synth_weights(d, **synth_preset("tiny")), True)
    out = wo.transcribe(w, d, speech_shaped_audio(7.0, 3), language="ja", temperature=0.0, condition_on_previous_text=False,
                        max_initial_timestamp=0.0, sample_len=24)
    assert out["language"] == "ja" and isinstance(out["segments"], list)
    for s in out["segments"]:
        assert s["end"] >= s["start"] >= 0.0 and {"id", "seek", "tokens", "avg_logprob", "no_speech_prob", "compression_ratio"} <= set(s)
    # empty audio -> no windows
    assert wo.transcribe(w, d, np.zeros(0, np.float32), language="ja")["segments"] == []
