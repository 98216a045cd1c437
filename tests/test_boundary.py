"""Plug-in surface contracts that need no GPU (the reference's own contract tests restated:
tests/test_anime_whisper.py:129-288, tests/test_speech_segmentation.py:170-250)."""
import inspect
import sys
from pathlib import Path

import numpy as np
import pytest

import whisperjav_b200
from whisperjav_b200 import hostlogic as H
from whisperjav_b200.asr import B200WhisperASR
from whisperjav_b200.audioio import compose_srt, read_wav_mono, write_wav_pcm16
from whisperjav_b200.generator import B200WhisperGenerator
from whisperjav_b200.segmenter import B200SpeechSegmenter


def test_generator_protocol_surface():
    g = B200WhisperGenerator(model_id="tiny", device="cuda", dtype="float16", no_repeat_ngram_size=0, max_new_tokens=444)
    for name in ("generate", "generate_batch", "load", "unload", "cleanup"):
        assert callable(getattr(g, name))
    assert g.is_loaded is False
    with pytest.raises(RuntimeError, match="before load"):
        g.generate(Path("x.wav"))
    with pytest.raises(RuntimeError, match="before load"):
        g.generate_batch([Path("x.wav")])
    sig = inspect.signature(g.generate_batch)
    assert list(sig.parameters)[:3] == ["audio_paths", "language", "contexts"]
    g.cleanup()  # unloading an unloaded generator is a no-op
    with pytest.raises(ValueError):
        B200WhisperGenerator(no_repeat_ngram_size=5)


def test_segmenter_surface_and_postprocess():
    s = B200SpeechSegmenter(threshold=0.4, chunk_threshold_s=2.5, max_group_duration_s=6.0, version="x", variant="y")
    assert s.name == "b200-vad" and isinstance(s.display_name, str) and s.get_supported_sample_rates() == [16000]
    probs = np.zeros(938, dtype=np.float32)
    probs[100:200] = 0.9   # 3.2 s .. 6.4 s
    probs[400:500] = 0.9   # 12.8 s .. 16.0 s
    r = s._postprocess(probs, 480000, 30.0, {}, 0.0)
    assert r.method == "b200-vad" and r.num_segments == 2 and r.num_groups == 2
    seg = r.segments[0]
    # padding: start - 11200 samples, end + 20800 samples (silero.py:286-297) on top of the hysteresis regions
    assert (seg.start_sample, seg.end_sample) == (100 * 512 - 11200, 200 * 512 + 20800)
    assert (r.segments[1].start_sample, r.segments[1].end_sample) == (400 * 512 - 11200, 500 * 512 + 20800)
    assert r.segments[1].start_sample >= r.segments[0].end_sample
    assert r.to_legacy_format()[0][0]["start"] == seg.start_sample
    t = B200SpeechSegmenter(threshold=0.5, style="ten", min_silence_duration_ms=100, chunk_threshold_s=1.0)
    rt = t._postprocess(probs, 480000, 30.0, {}, 0.0)
    assert rt.num_segments == 2 and abs(rt.segments[0].start_sec - (100 * 0.032 - 0.05)) < 1e-9
    assert abs(rt.segments[0].end_sec - (200 * 0.032 + 0.15)) < 1e-9 and "raw_start" in rt.segments[0].metadata
    empty = s._postprocess(np.zeros(100, np.float32), 51200, 3.2, {}, 0.0)
    assert empty.segments == [] and empty.groups == []
    s.cleanup()


def test_asr_wrapper_signature_matches_reference():
    params = inspect.signature(B200WhisperASR.__init__).parameters
    assert list(params)[1:] == ["model_config", "params", "task", "tracer"]
    for m in ("transcribe", "transcribe_to_srt", "reset_statistics", "get_filter_statistics", "get_last_vad_segments", "cleanup"):
        assert callable(getattr(B200WhisperASR, m))


def test_wav_and_srt_io(tmp_path):
    a = (0.5 * np.sin(np.arange(16000) * 0.05)).astype(np.float32)
    write_wav_pcm16(tmp_path / "a.wav", a)
    b, sr = read_wav_mono(tmp_path / "a.wav")
    assert sr == 16000 and len(b) == len(a) and np.abs(a - b).max() < 1e-4
    srt_text = compose_srt([{"start": 1.5, "end": 3.25, "text": "こんにちは"}, {"start": 3661.001, "end": 3662.0, "text": "x"}])
    assert "1\n00:00:01,500 --> 00:00:03,250\nこんにちは\n" in srt_text and "01:01:01,001 --> 01:01:02,000" in srt_text
    assert compose_srt([]) == ""


@pytest.mark.skipif(not Path("/root/reference/whisperjav").exists(), reason="reference tree not present (GPU box)")
def test_registration_with_reference_factories(monkeypatch):
    import types
    sys.path.insert(0, "/root/reference")
    for absent in ("librosa", "soundfile"):  # imported at module level by scene_detection_backends/utils.py, not installed here
        if absent not in sys.modules:
            monkeypatch.setitem(sys.modules, absent, types.ModuleType(absent))
    try:
        done = whisperjav_b200.register()
        assert done == {"speech_segmenter": True, "text_generator": True, "scene_detector": True}
        from whisperjav.modules.scene_detection_backends.base import SceneDetector
        from whisperjav.modules.scene_detection_backends.factory import SceneDetectorFactory
        det = SceneDetectorFactory.create("b200-auditok", max_duration=20.0, pass2_max_silence_s=0.5)
        assert isinstance(det, SceneDetector) and det.name == "b200-auditok"
        assert det._config.max_duration == 20.0 and det._config.pass2_max_duration == 19.0 and det._config.pass2_max_silence == 0.5
        assert SceneDetectorFactory.is_backend_available("b200-auditok") == (True, "")
        sil = SceneDetectorFactory.create("b200-silero", silero_threshold=0.1)
        assert isinstance(sil, SceneDetector) and sil.name == "b200-silero"
        assert sil._silero_config.max_duration == 420.0 and sil._silero_config.brute_force_chunk_s == 29.0 and sil._silero_config.silero_threshold == 0.1
        from whisperjav.modules.speech_segmentation import SpeechSegmenterFactory
        from whisperjav.modules.speech_segmentation.base import SpeechSegmenter
        from whisperjav.modules.subtitle_pipeline.generators.factory import TextGeneratorFactory
        from whisperjav.modules.subtitle_pipeline.protocols import TextGenerator
        seg = SpeechSegmenterFactory.create("b200-vad", config={"threshold": 0.4, "chunk_threshold_s": 2.5})
        assert isinstance(seg, SpeechSegmenter) and seg.chunk_threshold_s == 2.5
        ws = SpeechSegmenterFactory.create("b200-whisperseg", config={"threshold": 0.35, "max_group_duration_s": 6.0})
        assert isinstance(ws, SpeechSegmenter) and ws.name == "b200-whisperseg" and ws.max_speech_duration_s == 6.0
        gen = TextGeneratorFactory.create("b200-whisper", model_id="tiny", device="cuda", dtype="float16",
                                          no_repeat_ngram_size=0, max_new_tokens=444)
        assert isinstance(gen, TextGenerator)
    finally:
        sys.path.remove("/root/reference")


def test_whisperseg_surface_and_postprocess_match_reference_defaults():
    """Constructor keywords / defaults of whisperseg.py:80-141 and the probs -> segments -> groups chain on scripted frame
    probabilities (the state machine itself is pinned by the reference-generated KATs in test_host_kats.py)."""
    from whisperjav_b200.whisperseg import B200WhisperSegSegmenter
    s = B200WhisperSegSegmenter(version="x")
    assert (s.threshold, s.min_speech_duration_ms, s.min_silence_duration_ms, s.speech_pad_ms) == (0.35, 100, 100, 300)
    assert (s.chunk_threshold_s, s.max_group_duration_s, s.max_speech_duration_s) == (1.0, 29.0, 29.0)
    assert s.name == "b200-whisperseg" and s.get_supported_sample_rates() == [16000]
    p = np.zeros(1500, np.float32)
    p[100:200] = 0.9
    p[260:300] = 0.9
    r = s._postprocess(p, 30.0, 0.0)
    # whisperseg.py: p[100:200] = 0.9 -> one segment 1.700-4.300 s (2.0 - 0.3, 4.0 + 0.3); the second region is
    # 1.2 s later, so its padding is clipped half way to the first and they fall into one group (gap < 1.0 s)
    assert abs(r.segments[0].start_sec - 1.7) < 1e-9 and r.num_segments == 2 and r.num_groups == 1
    assert B200WhisperSegSegmenter(chunk_threshold_s=None, chunk_threshold=2.0).chunk_threshold_s == 2.0


@pytest.mark.skipif(not Path("/root/reference/whisperjav").exists(), reason="reference tree not present (GPU box)")
def test_reference_vad_grouped_framer_drives_the_b200_segmenters(monkeypatch):
    """The reference's own VadGroupedFramer (subtitle_pipeline/framers/vad_grouped.py:77-164) run with both B200 segmenters:
    factory -> segment() -> groups -> TemporalFrames.  The device stage is replaced by scripted probabilities (no GPU here); the
    frames must be exactly the groups of the reference's own state machines on those probabilities."""
    sys.path.insert(0, "/root/reference")
    try:
        whisperjav_b200.register()
        from whisperjav.modules.subtitle_pipeline.framers.vad_grouped import VadGroupedFramer
        from whisperjav_b200.segmenter import B200SpeechSegmenter
        from whisperjav_b200.whisperseg import B200WhisperSegSegmenter
        audio = np.zeros(480000, np.float32)

        class FakeVad:
            device = "cpu"

            def probs(self, a, ns):
                import torch
                p = torch.zeros(a.shape[0], (a.shape[1] + 511) // 512)
                p[:, 100:200] = 0.9
                p[:, 400:500] = 0.9
                return p
        monkeypatch.setattr(B200SpeechSegmenter, "_ensure_model", lambda self: FakeVad())
        fr = VadGroupedFramer(segmenter_backend="b200-vad", max_group_duration_s=6.0, chunk_threshold_s=2.5).frame(audio, 16000)
        assert fr.metadata["segmenter_backend"] == "b200-vad" and fr.metadata["total_segments"] == 2
        assert [(round(f.start, 3), round(f.end, 3)) for f in fr.frames] == [(2.5, 7.7), (12.1, 17.3)]

        def fake_probs(self, clips):
            p = np.zeros(1500, np.float32)
            p[100:200] = 0.9
            return [p for _ in clips]
        monkeypatch.setattr(B200WhisperSegSegmenter, "frame_probs", fake_probs)
        fr = VadGroupedFramer(segmenter_backend="b200-whisperseg").frame(audio, 16000)
        assert fr.metadata["segmenter_backend"] == "b200-whisperseg"
        assert [(round(f.start, 3), round(f.end, 3)) for f in fr.frames] == [(1.7, 4.3)]     # 2.0 - 0.3 .. 4.0 + 0.3
        assert fr.metadata["speech_regions"] == [[(1.7, 4.3)]]
    finally:
        sys.path.remove("/root/reference")
