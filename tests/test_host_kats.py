"""Host-side logic vs known answers produced by the reference's own code (tests/golden/make_host_kats.py),
plus the reference tests' own KATs restated (tests/test_vad_failover.py:22-31,
tests/test_speech_segmentation.py:144-166)."""
import json
from pathlib import Path

import pytest

from whisperjav_b200 import hostlogic as H

KATS = json.loads((Path(__file__).parent / "golden" / "reference_host_kats.json").read_text())


def test_group_by_gap_matches_reference():
    for c in KATS["group_segments"]:
        segs = [H.SpeechSegment(a, b) for a, b in c["segments"]]
        got = H.group_by_gap(segs, c["max_group_duration_s"], c["chunk_threshold_s"])
        assert [[(s.start_sec, s.end_sec) for s in g] for g in got] == [[tuple(x) for x in g] for g in c["groups"]]
    assert H.group_by_gap([]) == []


def test_silero_pad_clamp_group_matches_reference():
    for c in KATS["silero_pad_group"]:
        padded = H.pad_and_clamp(c["timestamps"], c["n_audio"])
        assert padded == [tuple(x) for x in c["segments"]]
        segs = [H.SpeechSegment(a / 16000, b / 16000, a, b) for a, b in padded]
        groups = H.group_by_gap(segs, c["max_group_duration_s"], c["chunk_threshold_s"])
        assert [[(s.start_sample, s.end_sample) for s in g] for g in groups] == [[tuple(x) for x in g] for g in c["groups"]]


def test_probability_state_machine_matches_reference():
    import numpy as np
    for c in KATS["prob_state_machine"]:
        p = np.array(c["probs"], dtype=np.float32)
        got = H.probs_to_regions(p, len(p) * 0.02, frame_ms=c["frame_ms"], threshold=c["threshold"],
                                 min_speech_duration_ms=c["min_speech_duration_ms"], min_silence_duration_ms=c["min_silence_duration_ms"],
                                 speech_pad_ms=c["speech_pad_ms"], max_speech_duration_s=c["max_speech_duration_s"])
        assert [(s.start_sec, s.end_sec) for s in got] == [tuple(x) for x in c["segments"]]


def test_survey_kat_whisperseg_example():
    """threshold 0.35, p[100:200] = 0.9 -> one segment 1.700-4.300 s (defaults of the backend)."""
    import numpy as np
    p = np.zeros(1500, dtype=np.float32)
    p[100:200] = 0.9
    got = H.probs_to_regions(p, 30.0, frame_ms=20.0, threshold=0.35, min_speech_duration_ms=250, min_silence_duration_ms=100,
                             speech_pad_ms=300, max_speech_duration_s=0.0)
    assert len(got) == 1 and abs(got[0].start_sec - 1.7) < 1e-9 and abs(got[0].end_sec - 4.3) < 1e-9


def test_vad_failover_matches_reference():
    for c in KATS["force_full_transcribe"]:
        assert H.vad_looks_broken(c["vad"], c["duration"]) == c["expect"], c
    # reference tests/test_vad_failover.py:22-31
    assert H.vad_looks_broken([], 600.0) is True
    assert H.vad_looks_broken([], 30.0) is False


def test_logprob_gate_matches_reference():
    for c in KATS["segment_filter"]:
        g = H.LogprobGate(True, c["thr"], c["margin"], c["nonverbal"])
        f, reason, eff = g.should_filter(c["lp"], c["dur"], c["text"])
        assert (f, reason) == (c["filter"], c["reason"]), c
        assert eff == c["eff"] or (eff is not None and abs(eff - c["eff"]) < 1e-12)
    assert H.LogprobGate(enabled=False, logprob_threshold=-1.0).should_filter(-5.0, 1.0, "x") == (False, None, None)


def test_legacy_format():
    # reference tests/test_speech_segmentation.py:144-166
    s = H.SpeechSegment(1.0, 2.0, 16000, 32000, metadata={"k": 1})
    r = H.SegmentationResult([s], [[s]], "b200", 3.0, {})
    assert r.to_legacy_format() == [[{"start": 16000, "end": 32000, "start_sec": 1.0, "end_sec": 2.0, "metadata": {"k": 1}}]]
    assert r.num_segments == 1 and r.num_groups == 1 and abs(r.speech_coverage_ratio - 1 / 3) < 1e-12


def test_ten_style_pipeline_matches_reference():
    for c in KATS["ten_pipeline"]:
        raw = H.flags_to_regions(c["flags"], c["probs"], 256 / 16000, c["duration"], c["min_speech_duration_ms"], c["max_speech_duration_s"])
        assert [(r["start"], r["end"], len(r["probs"])) for r in raw] == [tuple(x) for x in c["raw"]]
        final = H.ten_style_segments(c["flags"], c["probs"], c["duration"], min_speech_duration_ms=c["min_speech_duration_ms"],
                                     min_silence_duration_ms=c["min_silence_duration_ms"], max_speech_duration_s=c["max_speech_duration_s"],
                                     start_pad_ms=c["start_pad_ms"], end_pad_ms=c["end_pad_ms"])
        got = [(s.start_sec, s.end_sec, s.confidence, s.metadata["raw_start"], s.metadata["raw_end"]) for s in final]
        assert len(got) == len(c["final"])
        for g, r in zip(got, c["final"]):
            assert g[0] == r[0] and g[1] == r[1] and abs(g[2] - r[2]) < 1e-9 and g[3] == r[3] and g[4] == r[4]
