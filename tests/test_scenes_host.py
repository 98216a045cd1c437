"""Scene detector, host side (CPU): the product's tokenizer / two-pass driver against (i) the frame-list restatement of
auditok's StreamTokenizer in oracle/scene_oracle.py on random flag sequences and (ii) scenes produced by the reference's own
``AuditokSceneDetector.detect_scenes`` (tests/golden/make_scene_kats.py -> reference_scene_kats.json).  The energy provider here
is the oracle's numpy twin of the kernel; tests/test_gpu_scenes.py swaps in the kernel."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import scene_oracle as SO
from scene_cases import CASES, build_case
from whisperjav_b200 import scenes as SC

KATS = {k["name"]: k for k in json.loads((Path(__file__).parent / "golden" / "reference_scene_kats.json").read_text())}


def test_tokenizer_matches_stream_tokenizer_on_random_flags():
    rng = np.random.default_rng(5)
    for it in range(400):
        n = int(rng.integers(0, 300))
        p = float(rng.choice([0.1, 0.5, 0.8, 0.95]))
        run = int(rng.integers(1, 12))
        flags = np.repeat(rng.random(n // run + 1) < p, run)[:n].tolist()
        max_len = int(rng.integers(1, 60))
        min_len = int(rng.integers(1, max_len + 1))
        max_sil = int(rng.integers(0, max_len))
        drop, strict = bool(rng.integers(2)), bool(rng.integers(2))
        mode = (SO.StreamTokenizer.DROP_TRAILING_SILENCE if drop else 0) | (SO.StreamTokenizer.STRICT_MIN_LENGTH if strict else 0)
        ref = SO.StreamTokenizer(lambda f: f, min_len, max_len, max_sil, mode=mode).tokenize(flags)
        got = SC.tokenize_flags(flags, min_len, max_len, max_sil, drop, strict)
        assert got == [(s, len(d)) for d, s, _ in ref], (it, min_len, max_len, max_sil, drop, strict)
        for d, s, e in ref:
            assert e == s + len(d) - 1


def test_tokenizer_rejects_what_auditok_rejects():
    with pytest.raises(ValueError):
        SC.tokenize_flags([True], 3, 2, 0)
    with pytest.raises(ValueError):
        SC.tokenize_flags([True], 1, 4, 4)
    with pytest.raises(ValueError):
        SO.StreamTokenizer(lambda f: f, 1, 4, 4)


def test_window_rounding_rules():
    # auditok.core._duration_to_nb_windows: ceil for min_dur, floor(x + 1e-10) for max_dur / max_silence
    for d in (0.0, 0.05, 0.3, 0.94, 1.8, 28.0, 2700.0, 0.149999, 0.15):
        assert SC.windows_for(d, 0.05, True) == SO.duration_to_nb_windows(d, 0.05, math.ceil)
        assert SC.windows_for(d, 0.05, False) == SO.duration_to_nb_windows(d, 0.05, math.floor, 1e-10)
    assert SC.windows_for(0.3, 0.05, True) == 6 and SC.windows_for(0.94, 0.05, False) == 18 and SC.windows_for(1.8, 0.05, False) == 36


def test_energy_flags_are_upstreams_decision():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((200, 800)) * rng.choice([5, 20, 40, 80, 300], size=(200, 1))).astype(np.int16)
    x[0] = 0
    ss = np.array([int(np.sum(r.astype(np.int64) ** 2)) for r in x], dtype=np.uint64)
    for thr in (32, 38, 50):
        want = np.array([SO.calculate_energy(r) >= thr for r in x])
        assert np.array_equal(SC.energy_flags(ss, np.full(200, 800), thr), want)


def _numpy_energy(audio):
    return lambda regions, window: SO.window_sumsq(audio, regions, window)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_two_pass_reproduces_the_reference_driver(case):
    audio, sr = build_case(case)
    cfg = SC.config_from_kwargs(case.get("kwargs", {}))
    scenes, story, counters = SC.two_pass(cfg, len(audio), sr, _numpy_energy(audio))
    kat = KATS[case["name"]]
    assert [[s.start_sec, s.end_sec, s.detection_pass, s.metadata.get("split_method", "")] for s in scenes] == kat["scenes"]
    assert [[round(a, 3), round(b, 3)] for a, b in story] == kat["coarse"]
    assert sum(counters.values()) == len(kat["scenes"])


def test_oracle_driver_agrees_with_the_reference_driver():
    """the restated driver in the oracle (used as the CPU baseline) against the same known answers"""
    for case in CASES:
        audio, sr = build_case(case)
        kw = SC.config_from_kwargs(case.get("kwargs", {}))
        got, story = SO.detect_scenes(audio, sr, max_duration=kw.max_duration, min_duration=kw.min_duration, pass1_min_duration=kw.pass1_min_duration,
                                      pass1_max_duration=kw.pass1_max_duration, pass1_max_silence=kw.pass1_max_silence,
                                      pass1_energy_threshold=kw.pass1_energy_threshold, pass2_min_duration=kw.pass2_min_duration,
                                      pass2_max_duration=kw.pass2_max_duration, pass2_max_silence=kw.pass2_max_silence,
                                      pass2_energy_threshold=kw.pass2_energy_threshold, brute_force_fallback=kw.brute_force_fallback,
                                      brute_force_chunk_s=kw.brute_force_chunk_s, pad_edges_s=kw.pad_edges_s)
        kat = KATS[case["name"]]
        assert [[s, e, p, "brute_force" if m == "brute_force" else ""] for s, e, p, m in got] == kat["scenes"], case["name"]


def test_config_aliases_and_derived_defaults():
    c = SC.config_from_kwargs({"max_duration_s": 20, "max_silence": 1.1, "energy_threshold": 30})
    assert (c.max_duration, c.pass1_max_silence, c.pass1_energy_threshold, c.pass2_max_duration, c.brute_force_chunk_s) == (20.0, 1.1, 30, 19.0, 20.0)
    d = SC.SceneConfig()
    assert (d.pass2_max_duration, d.pass2_max_silence, d.pass2_energy_threshold, d.pass1_max_duration) == (28.0, 0.94, 38, 2700.0)


def test_detector_needs_a_gpu_and_registers_with_the_reference_factory():
    import torch
    det = SC.B200SceneDetector(max_duration=29.0)
    assert det.name == "b200-auditok" and callable(det.detect_scenes) and callable(det.cleanup)
    if not torch.cuda.is_available():
        from whisperjav_b200._lib import WjbError
        with pytest.raises(WjbError):
            det.detect(np.zeros(16000, dtype=np.float32), 16000)


# ---- Silero-style detector: energy pass 1 + VAD pass 2 ----------------------------------------------------------------------------

SILERO_KATS = {k["name"]: k for k in json.loads((Path(__file__).parent / "golden" / "reference_silero_scene_kats.json").read_text())}


class FakeVad:
    """stands in for vad.VadB200 with the same call shape: probs(audio [B, S], n_samples [B]) -> [B, windows]"""
    device = "cpu"

    def probs(self, audio, n_samples):
        import torch
        from scene_cases import fake_vad_probs
        rows = [fake_vad_probs(audio[i, : int(n_samples[i])].numpy()) for i in range(audio.shape[0])]
        out = torch.zeros(len(rows), (audio.shape[1] + 511) // 512)
        for i, r in enumerate(rows):
            out[i, : len(r)] = torch.from_numpy(r)
        return out


def test_silero_style_detector_reproduces_the_reference_driver():
    """SileroSceneDetector.detect_scenes of the reference (run with a fake VAD, tests/golden/make_silero_scene_kats.py) against
    B200SileroSceneDetector with the same fake VAD and the numpy twin of the energy kernel: config derivation, which chapters reach
    pass 2, scenes, the brute-force fall-through, the VAD-segment metadata."""
    from scene_cases import SILERO_CASES
    for case in SILERO_CASES:
        audio, sr = build_case(case)
        det = SC.B200SileroSceneDetector(vad=FakeVad(), **case.get("kwargs", {}))
        kat = SILERO_KATS[case["name"]]
        cfg = det._silero_config
        assert {"max_duration": cfg.max_duration, "pass2_max_duration": cfg.pass2_max_duration, "brute_force_chunk_s": cfg.brute_force_chunk_s,
                "min_duration": cfg.min_duration, "silero_threshold": cfg.silero_threshold, "assist_processing": cfg.assist_processing} == kat["config"]
        det._vad_segments = []
        scenes, story, counters = SC.two_pass(det._config, len(audio), sr, _numpy_energy(audio), det._fine_split(audio, sr))
        assert [[s.start_sec, s.end_sec, s.detection_pass, s.metadata.get("split_method", "")] for s in scenes] == kat["scenes"], case["name"]
        assert [[round(a, 3), round(b, 3)] for a, b in story] == kat["coarse"]
        assert (det._vad_segments or None) == kat["vad_segments"]
    assert SC.B200SileroSceneDetector(vad=FakeVad()).name == "b200-silero"


def test_hysteresis_takes_an_explicit_negative_threshold():
    from whisperjav_b200 import hostlogic as H
    p = np.array([0.0] * 10 + [0.5] * 30 + [0.12] * 30 + [0.5] * 30 + [0.0] * 40, dtype=np.float32)
    kw = dict(frame_ms=32.0, threshold=0.3, min_speech_duration_ms=100, min_silence_duration_ms=320, speech_pad_ms=0)
    one = H.probs_to_regions(p, len(p) * 0.032, **kw)                        # default offset threshold 0.15: 0.12 is silence -> two regions
    two = H.probs_to_regions(p, len(p) * 0.032, neg_threshold=0.1, **kw)     # 0.12 stays speech -> one region
    assert len(one) == 2 and len(two) == 1
