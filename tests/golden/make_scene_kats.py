"""Scene-detector known answers produced by *running the reference's own two-pass driver* in this container:

    PYTHONPATH=/root/reference:. python tests/golden/make_scene_kats.py

``whisperjav.modules.scene_detection_backends.auditok_backend.AuditokSceneDetector.detect_scenes`` (auditok_backend.py:229-524) is
executed unmodified; the two third-party packages it imports that are absent here are stood in for: ``auditok`` by
oracle/scene_oracle.py (the restated ``auditok.split``, so THAT function stays unpinned -- see its header) and ``soundfile`` / ``librosa`` by
empty stubs (``sf.write`` records nothing; no resampling happens on this path).  The audio is regenerated from seeds by the tests (whisperjav_b200.synth.film_audio + the
hand-built cases below), the JSON holds the reference's scenes / coarse boundaries per case and configuration.
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
sys.path.insert(0, "/root/reference")

from oracle import scene_oracle  # noqa: E402

auditok_stub = types.ModuleType("auditok")
auditok_stub.split = scene_oracle.split
sys.modules["auditok"] = auditok_stub
sf_stub = types.ModuleType("soundfile")
sf_stub.write = lambda *a, **k: None
sys.modules["soundfile"] = sf_stub
sys.modules["librosa"] = types.ModuleType("librosa")  # imported by utils.py for resampling; never called on this path

from whisperjav.modules.scene_detection_backends import auditok_backend as ab  # noqa: E402
import whisperjav.modules.scene_detection_backends.utils as ref_utils  # noqa: E402

ref_utils.sf = sf_stub
sys.path.insert(0, str(ROOT / "tests"))
from scene_cases import CASES, build_case  # noqa: E402

out = []
for case in CASES:
    audio, sr = build_case(case)
    det = ab.AuditokSceneDetector(**case.get("kwargs", {}))
    det._load_audio = lambda path, _a=audio, _sr=sr: (_a, _sr)
    with tempfile.TemporaryDirectory() as td:
        res = det.detect_scenes(Path("unused.wav"), Path(td), "kat")
    out.append({
        "name": case["name"],
        "scenes": [[s.start_sec, s.end_sec, s.detection_pass, s.metadata.get("split_method", "")] for s in res.scenes],
        "coarse": [[c["start_time_seconds"], c["end_time_seconds"]] for c in res.coarse_boundaries],
        "files": [s.scene_path.name for s in res.scenes],
        "audio_duration_sec": res.audio_duration_sec,
    })
    print(case["name"], len(res.scenes), "scenes,", len(res.coarse_boundaries), "story lines,",
          sum(1 for s in res.scenes if s.metadata.get("split_method") == "brute_force"), "brute-force")
(HERE / "reference_scene_kats.json").write_text(json.dumps(out, indent=0))
