"""Boundary classes end to end on the GPU: VAD kernel vs its oracle, segmenter, ASR wrapper -> SRT,
TextGenerator batch."""
import json

import numpy as np
import pytest
import torch

from oracle.vad_oracle import vad_probs
from whisperjav_b200.audioio import write_wav_pcm16
from oracle import whisper_oracle as wo
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_preset, synth_weights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["energy", "random"])
def test_vad_kernel_matches_oracle(kind, diag_dir):
    from whisperjav_b200.vad import VadB200, energy_vad_weights, synth_vad_weights
    sd = energy_vad_weights(5) if kind == "energy" else synth_vad_weights(9)
    vad = VadB200(sd)
    clips = [speech_shaped_audio(s, 300 + i) for i, s in enumerate([12.0, 3.3, 0.02])]
    S = max(len(c) for c in clips)
    audio = torch.zeros(len(clips), S)
    for i, c in enumerate(clips):
        audio[i, : len(c)] = torch.from_numpy(c)
    ns = torch.tensor([len(c) for c in clips], dtype=torch.int32)
    got = vad.probs(audio.cuda(), ns.cuda()).cpu().numpy()
    for i, c in enumerate(clips):
        ref = vad_probs(sd, c)
        err = np.abs(got[i, : len(ref)] - ref).max()
        (diag_dir / f"vad_{kind}_{i}.json").write_text(json.dumps({"err": float(err), "n": len(ref)}))
        assert err <= 2e-3, (i, err)  # fp32 both sides; differences are summation order through the recurrent state
        assert np.all(got[i, len(ref):] == 0)


def test_segmenter_end_to_end():
    from whisperjav_b200.segmenter import B200SpeechSegmenter
    seg = B200SpeechSegmenter(threshold=0.5, min_silence_duration_ms=300, chunk_threshold_s=2.5, max_group_duration_s=6.0)
    a = speech_shaped_audio(30.0, 1000)
    r = seg.segment(a, sample_rate=16000)
    assert r.method == "b200-vad" and r.num_segments >= 3 and 0.3 < r.speech_coverage_ratio <= 1.0
    for g in r.groups:
        assert g[-1].end_sec - g[0].start_sec <= 6.0 + 1e-6 or len(g) == 1
    silent = seg.segment(np.zeros(16000 * 5, np.float32), sample_rate=16000)
    assert silent.num_segments == 0  # reference structural test: silence -> 0 segments
    rs = seg.segment_batch([a, a[:80000]])
    assert [s.start_sample for s in rs[0].segments] == [s.start_sample for s in r.segments]


def test_asr_wrapper_to_srt(tmp_path):
    """B200WhisperASR.transcribe / transcribe_to_srt against the oracle chain: the same VAD groups, each group through the
    oracle's transcribe() with the parameters the wrapper resolved, then the wrapper's own segment post-filter (host logic pinned
    by reference KATs).  Groups are short (<= 6 s: a few dozen tokens), so most must agree exactly -- text, times, avg_logprob;
    a group may differ only by a fp16 near-tie flip, and at most one may."""
    from whisperjav_b200.asr import B200WhisperASR
    from whisperjav_b200.audioio import read_wav_mono
    a = speech_shaped_audio(20.0, 77)
    wav = tmp_path / "scene_0001.wav"
    write_wav_pcm16(wav, a)
    params = {"decoder": {"task": "transcribe", "language": "ja", "suppress_blank": True, "without_timestamps": False, "max_initial_timestamp": 0.0},
              "provider": {"temperature": [0.0], "compression_ratio_threshold": 2.4, "logprob_threshold": -5.0, "no_speech_threshold": 0.71,
                           "condition_on_previous_text": False, "word_timestamps": False, "fp16": True, "logprob_margin": 0.0},
              "vad": {"threshold": 0.5, "chunk_threshold_s": 2.5, "max_group_duration_s": 6.0},
              "speech_segmenter": {"backend": "b200-vad"}}
    w = synth_weights(DIMS["tiny"], **synth_preset("tiny"))
    asr = B200WhisperASR({"model_name": "tiny", "device": "cuda", "state_dict": w}, params, "transcribe")
    out = asr.transcribe_to_srt(wav, tmp_path / "out" / "scene_0001.srt")
    assert out.exists() and out.stat().st_size > 0
    text = out.read_text(encoding="utf-8")
    assert "-->" in text and text.startswith("1\n")
    res = asr.transcribe(wav)
    assert res["language"] == "ja" and len(res["segments"]) >= 1
    assert all(0.0 <= s["start"] <= s["end"] <= 20.0 + 30.0 for s in res["segments"])  # timestamp tokens are not clamped to the content
    assert set(asr.get_filter_statistics()) == {"logprob_filtered", "nonverbal_filtered"}
    assert all(set(v) == {"start_sec", "end_sec"} for v in asr.get_last_vad_segments())
    # ---- oracle chain
    audio, sr = read_wav_mono(wav)
    groups = asr._external_segmenter.segment(audio, sample_rate=sr).to_legacy_format()
    assert len(groups) >= 2
    pw = wo.prepare_weights(w, True)
    kw = {k: v for k, v in asr._prepare_whisper_params().items() if k not in ("verbose", "fp16", "word_timestamps")}
    kw["temperature"] = 0.0
    expect, per_group = [], []
    for g in groups:
        s0, e0 = g[0]["start_sec"], g[-1]["end_sec"]
        ref = wo.transcribe(pw, DIMS["tiny"], audio[int(s0 * sr): int(e0 * sr)], **kw)
        segs = asr._process_segments(ref["segments"], s0)
        per_group.append((s0, e0, segs))
        expect.extend(segs)
    same_groups = 0
    for s0, e0, segs in per_group:
        mine = [s for s in res["segments"] if s0 - 1e-6 <= s["start"] and s["start"] < e0 + 30.0 and any(abs(s["start"] - x["start"]) < 1e-6 for x in segs)]
        if len(mine) == len(segs) and all(x["text"] == y["text"] and abs(x["end"] - y["end"]) < 1e-6 and abs(x["avg_logprob"] - y["avg_logprob"]) <= 2e-2
                                          for x, y in zip(mine, segs)):
            same_groups += 1
    assert same_groups >= len(per_group) - 1, (same_groups, len(per_group))
    if same_groups == len(per_group):
        assert [s["text"] for s in res["segments"]] == [s["text"] for s in expect]
    asr.cleanup()


def test_generator_batch_matches_single(tmp_path):
    from whisperjav_b200.generator import B200WhisperGenerator
    g = B200WhisperGenerator(model_id="tiny", device="cuda", state_dict=synth_weights(DIMS["tiny"], **synth_preset("tiny")), max_new_tokens=64)
    g.load()
    paths = []
    for i, s in enumerate([4.0, 2.5, 5.0]):
        p = tmp_path / f"f{i}.wav"
        write_wav_pcm16(p, speech_shaped_audio(s, 900 + i))
        paths.append(p)
    batch = g.generate_batch(paths, language="ja", contexts=None, audio_durations=[4.0, 2.5, 5.0])
    assert len(batch) == 3 and all(r.language == "ja" and r.metadata["generator"] == "b200-whisper" for r in batch)
    single = g.generate(paths[1])
    assert single.metadata["tokens"] == batch[1].metadata["tokens"]  # batching does not change a row's result
    assert len({tuple(r.metadata["tokens"]) for r in batch}) == 3
    g.unload()
    assert g.is_loaded is False


def test_whisperseg_class_segmenter(diag_dir):
    """The WhisperSeg-class gate end to end on the device: 30 s chunks -> HF-semantics 80-mel -> Whisper-base-shaped encoder ->
    frame head -> probs, checked against the CPU oracle's encoder (sim fp16) + the same head; then the reference's state machine."""
    from oracle import whisper_oracle as wo
    from whisperjav_b200.whisperseg import B200WhisperSegSegmenter
    seg = B200WhisperSegSegmenter(threshold=0.35, max_group_duration_s=6.0, chunk_threshold_s=1.0)
    clips = [speech_shaped_audio(41.0, 31), speech_shaped_audio(7.0, 32)]
    probs = seg.frame_probs(clips)
    assert [len(p) for p in probs] == [3000, 1500] and all(np.all((p >= 0) & (p <= 1)) for p in probs)
    # oracle: HF-semantics mel (raw audio padded to 30 s) -> encoder -> head
    m = seg._model
    dims = m.dims
    from whisperjav_b200.synth import synth_weights
    w = wo.prepare_weights(synth_weights(dims, seed=seg._seed), True)
    hw, hb = seg._head
    a = np.zeros(480000, np.float32)
    a[: len(clips[1])] = clips[1]
    mel = wo.log_mel_spectrogram(a, 80)[None]
    xa = wo.encoder_forward(w, dims, mel.half().float(), True)
    ref = torch.sigmoid(xa[0] @ hw.float().cpu() + hb).numpy()
    err = float(np.abs(probs[1] - ref).max())
    (diag_dir / "whisperseg.json").write_text(json.dumps({"err": err}))
    assert err <= 2e-2, err
    res = seg.segment_batch(clips)
    assert res[0].method == "b200-whisperseg" and abs(res[0].audio_duration_sec - 41.0) < 1e-6
    for r in res:
        for g in r.groups:
            assert g[-1].end_sec - g[0].start_sec <= 6.0 + 1e-6 or len(g) == 1
    assert seg.segment(np.zeros(0, np.float32)).segments == []
    seg.cleanup()
