"""Known answers for the Silero-style scene detector from the reference's own ``SileroSceneDetector.detect_scenes``
(silero_backend.py:52-299, inheriting the two-pass driver of auditok_backend.py):

    PYTHONPATH=/root/reference:. python tests/golden/make_silero_scene_kats.py

Stand-ins for the absent packages: ``auditok`` = oracle/scene_oracle.py (pass 1), ``soundfile`` / ``librosa`` = empty stubs, and
``silero_vad`` = a module whose ``get_speech_timestamps`` turns the FAKE window probabilities of tests/scene_cases.py::fake_vad_probs into
timestamps with this repo's hysteresis (hostlogic.probs_to_regions, explicit neg_threshold) and silero-vad's 0.1 s ``return_seconds``
rounding.  The VAD itself is therefore NOT what these vectors pin; they pin everything around it: the Silero config derivation
(420 s scenes, re-derived pass-2 limit, 29 s brute-force chunk), which chapters go to pass 2, how its regions become scenes, the
fall-through to the brute-force split, and the ``vad_segments`` metadata.
"""
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, "/root/reference")

from oracle import scene_oracle  # noqa: E402
from scene_cases import SILERO_CASES, build_case, fake_vad_probs  # noqa: E402
from whisperjav_b200 import hostlogic as H  # noqa: E402

for name in ("soundfile", "librosa", "auditok", "silero_vad"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["auditok"].split = scene_oracle.split
sys.modules["soundfile"].write = lambda *a, **k: None


def get_speech_timestamps(audio, model, sampling_rate=16000, threshold=0.5, neg_threshold=None, min_silence_duration_ms=100,
                          min_speech_duration_ms=250, max_speech_duration_s=float("inf"), min_silence_at_max_speech=98, speech_pad_ms=30,
                          return_seconds=False, **kw):
    x = np.asarray(audio, dtype=np.float32)
    p = fake_vad_probs(x)
    dur = len(x) / sampling_rate
    regs = H.probs_to_regions(p, dur, frame_ms=32.0, threshold=threshold, neg_threshold=neg_threshold, min_speech_duration_ms=min_speech_duration_ms,
                              min_silence_duration_ms=min_silence_duration_ms, speech_pad_ms=speech_pad_ms, max_speech_duration_s=max_speech_duration_s)
    assert return_seconds
    return [{"start": max(round(r.start_sample / sampling_rate, 1), 0), "end": min(round(r.end_sample / sampling_rate, 1), dur)} for r in regs]


sys.modules["silero_vad"].load_silero_vad = lambda *a, **k: object()
sys.modules["silero_vad"].get_speech_timestamps = get_speech_timestamps

import whisperjav.modules.scene_detection_backends.utils as ref_utils  # noqa: E402
from whisperjav.modules.scene_detection_backends import silero_backend as sb  # noqa: E402

ref_utils.sf = sys.modules["soundfile"]
out = []
for case in SILERO_CASES:
    audio, sr = build_case(case)
    det = sb.SileroSceneDetector(**case.get("kwargs", {}))
    det._load_audio = lambda path, _a=audio, _sr=sr: (_a, _sr)
    with tempfile.TemporaryDirectory() as td:
        res = det.detect_scenes(Path("unused.wav"), Path(td), "kat")
    cfg = det._silero_config
    out.append({"name": case["name"],
                "scenes": [[s.start_sec, s.end_sec, s.detection_pass, s.metadata.get("split_method", "")] for s in res.scenes],
                "coarse": [[c["start_time_seconds"], c["end_time_seconds"]] for c in res.coarse_boundaries],
                "vad_segments": res.vad_segments,
                "config": {"max_duration": cfg.max_duration, "pass2_max_duration": cfg.pass2_max_duration, "brute_force_chunk_s": cfg.brute_force_chunk_s,
                           "min_duration": cfg.min_duration, "silero_threshold": cfg.silero_threshold, "assist_processing": cfg.assist_processing}})
    print(case["name"], len(res.scenes), "scenes,", len(res.coarse_boundaries), "story lines,", len(res.vad_segments or []), "vad segments,",
          sum(1 for s in res.scenes if s.metadata.get("split_method") == "brute_force"), "brute-force")
(HERE / "reference_silero_scene_kats.json").write_text(json.dumps(out, indent=0))
