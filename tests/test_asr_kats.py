"""B1 -- the ASR wrapper's host logic against known answers produced by the reference's own ``WhisperProASR``
(tests/golden/make_asr_kats.py -> reference_asr_kats.json): constructor parameter unpacking (which keys reach ``transcribe()``,
thresholds, the post-model gate default), ``_prepare_whisper_params`` and ``_process_segments`` (suppression lists, the logprob /
non-verbal gate, offsets, statistics).  CPU only: the model loader is replaced by a stub."""
import json
from pathlib import Path

import pytest

from whisperjav_b200 import asr as A
from whisperjav_b200 import model as M

KATS = json.loads((Path(__file__).parent / "golden" / "reference_asr_kats.json").read_text())


@pytest.fixture()
def no_model(monkeypatch):
    monkeypatch.setattr(M, "load_model", lambda *a, **k: object())


@pytest.mark.parametrize("case", KATS, ids=[f"case{i}" for i in range(len(KATS))])
def test_wrapper_host_logic_matches_reference(case, no_model):
    params = json.loads(json.dumps(case["params"]))
    params["speech_segmenter"] = {"backend": "b200-vad"}   # the reference run used its "none" backend; the segmenter plays no part here
    asr = A.B200WhisperASR({"model_name": "large-v2", "device": "cuda"}, params, "transcribe")
    assert asr.logprob_threshold == case["logprob_threshold"] and asr.logprob_margin == case["logprob_margin"]
    assert asr.post_model_filter_enabled == case["post_model_filter_enabled"] and asr.drop_nonverbal_vocals == case["drop_nonverbal_vocals"]
    assert asr.suppress_low == case["suppress_low"] and asr.suppress_high == case["suppress_high"]
    assert asr.whisper_params == case["whisper_params"]
    prepared = {k: (list(v) if isinstance(v, tuple) else v) for k, v in asr._prepare_whisper_params().items()}
    assert prepared == case["prepared"]
    out = asr._process_segments([dict(s) for s in case["segments"]], case["start_sec"])
    assert out == case["out"]
    assert asr.get_filter_statistics() == case["stats"]
    asr.reset_statistics()
    assert asr.get_filter_statistics() == {"logprob_filtered": 0, "nonverbal_filtered": 0}
