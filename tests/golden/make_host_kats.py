"""Generate host-side known-answer vectors by *running the reference's own code* in this container
(/root/reference is importable for these modules; it does not exist on the GPU box, hence the fixture).

    PYTHONPATH=/root/reference:. python tests/golden/make_host_kats.py

Covers: group_segments (speech_segmentation/backends/ten.py:31-73), Silero post-VAD padding / clamp /
grouping (backends/silero.py:286-297,325-361, driven with a fake model exactly like the reference's
tests/test_vad_threshold_padding_e2e.py:403-437), the Silero-compatible probability state machine
(backends/whisperseg.py:419-571), should_force_full_transcribe (modules/vad_failover.py:26-57) and
SegmentFilterHelper.should_filter (modules/segment_filters.py:80-103).
"""
import json
import sys
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference")

from whisperjav.modules.segment_filters import SegmentFilterConfig, SegmentFilterHelper  # noqa: E402
from whisperjav.modules.speech_segmentation.backends.silero import SileroSpeechSegmenter  # noqa: E402
from whisperjav.modules.speech_segmentation.backends.ten import group_segments  # noqa: E402
from whisperjav.modules.speech_segmentation.backends.whisperseg import WhisperSegSpeechSegmenter  # noqa: E402
from whisperjav.modules.speech_segmentation.base import SpeechSegment  # noqa: E402
from whisperjav.modules.vad_failover import should_force_full_transcribe  # noqa: E402

rng = np.random.default_rng(20260922)
out = {}

# ---- group_segments
cases = []
for _ in range(12):
    n = int(rng.integers(0, 14))
    t, segs = 0.0, []
    for _ in range(n):
        t += float(rng.choice([0.05, 0.3, 0.9, 1.1, 2.6, 5.0]))
        d = float(rng.uniform(0.2, 7.0))
        segs.append((round(t, 3), round(t + d, 3)))
        t += d
    mg = float(rng.choice([5.0, 6.0, 29.0]))
    ct = float(rng.choice([0.5, 1.0, 2.5, 4.0]))
    groups = group_segments([SpeechSegment(a, b) for a, b in segs], max_group_duration_s=mg, chunk_threshold_s=ct)
    cases.append({"segments": segs, "max_group_duration_s": mg, "chunk_threshold_s": ct,
                  "groups": [[(s.start_sec, s.end_sec) for s in g] for g in groups]})
out["group_segments"] = cases

# ---- Silero padding / clamp / grouping with a fake model
cases = []
for _ in range(10):
    n_audio = int(rng.integers(16000, 16000 * 40))
    n = int(rng.integers(1, 8))
    pts = np.sort(rng.integers(0, n_audio, size=2 * n))
    ts = [{"start": int(pts[2 * i]), "end": int(pts[2 * i + 1])} for i in range(n) if pts[2 * i + 1] > pts[2 * i]]
    kw = dict(chunk_threshold_s=float(rng.choice([0.5, 2.5, 4.0])), max_group_duration_s=float(rng.choice([6.0, 29.0])))
    seg = SileroSpeechSegmenter(version="v4.0", **kw)
    seg._model = MagicMock()
    seg._get_speech_timestamps = lambda *a, _ts=ts, **k: [dict(x) for x in _ts]
    res = seg.segment(np.zeros(n_audio, dtype=np.float32), sample_rate=16000)
    cases.append({"n_audio": n_audio, "timestamps": ts, **kw,
                  "segments": [(s.start_sample, s.end_sample) for s in res.segments],
                  "groups": [[(s.start_sample, s.end_sample) for s in g] for g in res.groups]})
out["silero_pad_group"] = cases

# ---- probability state machine (20 ms frames)
cases = []
for k in range(10):
    T = int(rng.integers(50, 1500))
    p = np.clip(rng.normal(0.15, 0.1, T), 0, 1)
    for _ in range(int(rng.integers(0, 6))):
        a = int(rng.integers(0, T))
        b = min(T, a + int(rng.integers(3, 400)))
        p[a:b] = np.clip(rng.normal(0.8, 0.15, b - a), 0, 1)
    kw = dict(threshold=float(rng.choice([0.25, 0.35, 0.5])), min_speech_duration_ms=int(rng.choice([81, 150, 250])),
              min_silence_duration_ms=int(rng.choice([100, 300])), speech_pad_ms=int(rng.choice([0, 30, 300])),
              max_speech_duration_s=float(rng.choice([3.0, 10.0, 0.0])))
    seg = WhisperSegSpeechSegmenter(**kw)
    res = seg._probs_to_segments(p.astype(np.float32), T * 0.02)
    cases.append({"probs": [round(float(x), 4) for x in p], **kw, "frame_ms": float(seg._frame_duration_ms),
                  "segments": [(s.start_sec, s.end_sec) for s in res]})
    # the fixture stores 4-decimal probs: regenerate the answer from exactly those values
    p4 = np.array(cases[-1]["probs"], dtype=np.float32)
    res = seg._probs_to_segments(p4, T * 0.02)
    cases[-1]["segments"] = [(s.start_sec, s.end_sec) for s in res]
out["prob_state_machine"] = cases

# ---- failover
cases = []
for dur in (0.0, 60.0, 119.9, 120.0, 300.0, 480.0, 600.0):
    for segs in ([], [[{"start_sec": 1.0, "end_sec": 1.5}]], [[{"start_sec": 1.0, "end_sec": 9.0}], [{"start_sec": 20.0, "end_sec": 21.0}]],
                 [[{"start_sec": float(i), "end_sec": float(i) + 0.8} for i in range(0, 50, 2)]]):
        cases.append({"duration": dur, "vad": segs, "expect": bool(should_force_full_transcribe(segs, dur))})
out["force_full_transcribe"] = cases

# ---- segment filter
cases = []
for thr in (None, -1.0, -1.55):
    for margin in (0.0, 0.2):
        for nonverbal in (False, True):
            h = SegmentFilterHelper(SegmentFilterConfig(enabled=True, logprob_threshold=thr, logprob_margin=margin, drop_nonverbal_vocals=nonverbal))
            for lp in (-0.5, -1.1, -1.3, -1.7):
                for dur in (0.5, 1.6, 3.0):
                    for text in ("こんにちは", "♪♪", "ああ", "[music]", "mmm"):
                        f, reason, eff = h.should_filter(lp, dur, text)
                        cases.append({"thr": thr, "margin": margin, "nonverbal": nonverbal, "lp": lp, "dur": dur, "text": text,
                                      "filter": bool(f), "reason": reason, "eff": eff})
out["segment_filter"] = cases
(HERE / "reference_host_kats.json").write_text(json.dumps(out, ensure_ascii=False))
print({k: len(v) for k, v in out.items()})

# ---- TEN frame-flag pipeline (ten.py:280-515), appended fixture
from whisperjav.modules.speech_segmentation.backends.ten import TenSpeechSegmenter  # noqa: E402

rng2 = np.random.default_rng(777)
cases = []
for k in range(12):
    n = int(rng2.integers(20, 3000))
    flags = np.zeros(n, dtype=int)
    probs = np.clip(rng2.normal(0.1, 0.05, n), 0, 1)
    for _ in range(int(rng2.integers(0, 8))):
        a = int(rng2.integers(0, n))
        b = min(n, a + int(rng2.integers(2, 1200)))
        flags[a:b] = 1
        probs[a:b] = np.clip(0.7 + 0.25 * np.sin(np.arange(b - a) / rng2.uniform(3, 40)) + rng2.normal(0, 0.03, b - a), 0, 1)
    kw = dict(min_speech_duration_ms=int(rng2.choice([81, 200])), min_silence_duration_ms=int(rng2.choice([0, 100, 300])),
              max_speech_duration_s=float(rng2.choice([3.0, 10.0])), start_pad_ms=int(rng2.choice([0, 50])), end_pad_ms=int(rng2.choice([0, 150])))
    seg = TenSpeechSegmenter(**kw)
    probs4 = [round(float(x), 4) for x in probs]
    dur = n * 256 / 16000 - float(rng2.uniform(0, 0.01))
    raw = seg._flags_to_segments(flags.tolist(), probs4, 16000, dur)
    merged = seg._merge_by_silence(raw)
    padded = seg._apply_padding(merged, dur)
    final = seg._split_long_segments(padded)
    cases.append({"flags": flags.tolist(), "probs": probs4, "duration": dur, **kw,
                  "raw": [(r["start"], r["end"], len(r["probs"])) for r in raw],
                  "final": [(s.start_sec, s.end_sec, s.confidence, s.metadata["raw_start"], s.metadata["raw_end"]) for s in final]})
out["ten_pipeline"] = cases
(HERE / "reference_host_kats.json").write_text(json.dumps(out, ensure_ascii=False))
print({k: len(v) for k, v in out.items()})
