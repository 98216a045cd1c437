"""ASR-wrapper known answers produced by running the reference's own ``WhisperProASR`` (whisper_pro_asr.py:32-135 constructor,
:456-503 ``_process_segments``, ``_prepare_whisper_params``) in this container:

    PYTHONPATH=/root/reference:. python tests/golden/make_asr_kats.py

The packages it imports that are absent here (``whisper``, ``soundfile``, ``srt``) are stood in for by empty stubs: the constructor
only calls ``whisper.load_model`` (stubbed) and the methods exercised are pure host logic.  The segmenter backend is the reference's
own "none".  Output: tests/golden/reference_asr_kats.json.
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, "/root/reference")
for name in ("whisper", "soundfile", "srt", "librosa"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["whisper"].load_model = lambda *a, **k: object()

from whisperjav.modules.whisper_pro_asr import WhisperProASR  # noqa: E402

rng = np.random.default_rng(20260923)
TEXTS = ["こんにちは", "ありがとうございます", "ご視聴ありがとうございました", "字幕作成者 田中", "Thank you for watching", "視聴してね", "あっ", "んっ…", "はぁ",
         "提供は", "  ", "", "そうですね、わかりました", "Thanks for coming", "えっと", "スポンサーの皆様", "うん", "♪", "ああああああ"]
cases = []
for ci in range(10):
    decoder = {"task": "transcribe", "language": "ja", "beam_size": int(rng.choice([1, 2, 3])), "patience": 1.2,
               "logprob_threshold": [None, -1.0, -0.8, -1.3][ci % 4], "no_speech_threshold": 0.7}
    provider = {"temperature": [0.0, 0.17] if ci % 2 else 0.0, "fp16": True, "logprob_margin": [0.0, 0.2, None][ci % 3],
                "drop_nonverbal_vocals": bool(ci % 2), "post_model_filter_enabled": [None, True, False][ci % 3], "word_timestamps": True}
    params = {"decoder": decoder, "vad": {"threshold": 0.3, "chunk_threshold": 2.5}, "provider": provider, "speech_segmenter": {"backend": "none"}}
    asr = WhisperProASR({"model_name": "large-v2", "device": "cpu"}, params, "transcribe")
    segs = []
    for _ in range(40):
        a = float(np.round(rng.uniform(0, 25), 2))
        segs.append({"start": a, "end": float(np.round(a + rng.choice([0.0, 0.3, 0.8, 1.5, 2.5, 6.0]), 2)), "text": str(rng.choice(TEXTS)),
                     "avg_logprob": float(np.round(rng.uniform(-2.0, -0.05), 3))})
    start_sec = float(np.round(rng.uniform(0, 3000), 3))
    out = asr._process_segments([dict(s) for s in segs], start_sec)
    cases.append({"params": params, "segments": segs, "start_sec": start_sec, "out": out, "stats": dict(asr._filter_statistics),
                  "whisper_params": asr.whisper_params, "prepared": {k: (list(v) if isinstance(v, tuple) else v) for k, v in asr._prepare_whisper_params().items()},
                  "logprob_threshold": asr.logprob_threshold, "logprob_margin": asr.logprob_margin,
                  "post_model_filter_enabled": asr.post_model_filter_enabled, "drop_nonverbal_vocals": asr.drop_nonverbal_vocals,
                  "suppress_low": asr.suppress_low, "suppress_high": asr.suppress_high})
(HERE / "reference_asr_kats.json").write_text(json.dumps(cases, ensure_ascii=False, indent=0))
print(len(cases), "cases;", sum(len(c["out"]) for c in cases), "segments kept of", sum(len(c["segments"]) for c in cases),
      "; logprob-filtered", sum(c["stats"]["logprob_filtered"] for c in cases), "nonverbal-filtered", sum(c["stats"]["nonverbal_filtered"] for c in cases))
