"""Scene detector on the GPU: csrc/scene.cu (through the C-ABI) against the oracle's numpy twin -- exact integer equality -- and
the whole detector against the scenes the reference's own driver produced (tests/golden/reference_scene_kats.json)."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import scene_oracle as SO
from scene_cases import CASES, build_case
from whisperjav_b200 import scenes as SC
from whisperjav_b200.audioio import read_wav_mono, write_wav_pcm16

pytestmark = pytest.mark.gpu
KATS = {k["name"]: k for k in json.loads((Path(__file__).parent / "golden" / "reference_scene_kats.json").read_text())}


def test_window_energy_is_exact_incl_misaligned_regions_and_short_tails():
    rng = np.random.default_rng(1)
    audio = (rng.standard_normal(16000 * 90) * rng.choice([1e-4, 3e-3, 0.1, 0.5], size=16000 * 90)).astype(np.float32)
    audio[5000:5100] = [1.0, -1.0] * 50                      # full scale: 32767^2 per sample
    audio[7000:7004] = [1.00002, -1.00003, 0.99999, -0.5]   # just past full scale (int16 wrap of the reference's cast is kept)
    det = SC.B200SceneDetector()
    energy = det._energy_provider(audio)
    regions = [(0, len(audio)), (1, 12345), (3, 800), (16002, 799), (40000, 1), (100001, 160000 + 37), (len(audio) - 801, 801)]
    for window in (800, 1102, 64):
        got = energy(regions, window)
        want = SO.window_sumsq(audio, regions, window)
        for g, w, r in zip(got, want, regions):
            assert g.dtype == np.uint64 and np.array_equal(g, w), (window, r)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_detector_reproduces_the_reference_driver(case):
    audio, sr = build_case(case)
    det = SC.B200SceneDetector(**case.get("kwargs", {}))
    scenes, story, _ = det.detect(audio, sr)
    kat = KATS[case["name"]]
    assert [[s.start_sec, s.end_sec, s.detection_pass, s.metadata.get("split_method", "")] for s in scenes] == kat["scenes"]
    assert [[round(a, 3), round(b, 3)] for a, b in story] == kat["coarse"]
    # a tensor already on the device gives the same answer (the stream pipeline's hand-off)
    scenes2, _, _ = det.detect(torch.from_numpy(audio).cuda(), sr)
    assert [(s.start_sec, s.end_sec) for s in scenes2] == [(s.start_sec, s.end_sec) for s in scenes]


def test_detect_scenes_protocol_writes_scene_wavs(tmp_path):
    case = CASES[0]
    audio, sr = build_case(case)
    src = tmp_path / "film.wav"
    write_wav_pcm16(src, audio, sr)
    det = SC.B200SceneDetector()
    res = det.detect_scenes(src, tmp_path / "scenes", "film")
    assert res.method == "b200-auditok" and abs(res.audio_duration_sec - len(audio) / sr) < 1e-9
    assert len(res.scenes) >= 10 and len(res.coarse_boundaries) >= 2
    for i, s in enumerate(res.scenes):
        assert s.scene_path.name == f"film_scene_{i:04d}.wav"
        data, sr2 = read_wav_mono(s.scene_path)
        assert sr2 == sr and len(data) == int(s.end_sec * sr) - int(s.start_sec * sr)
        assert 0.2 <= s.end_sec - s.start_sec <= 29.0 + 1e-9
    assert all(a.end_sec <= b.start_sec + 1e-9 for a, b in zip(res.scenes, res.scenes[1:]))
    det.cleanup()
