"""Host-side decision logic either side of the device path, with the reference's exact semantics
(pinned by known-answer vectors generated from the reference's own code, tests/golden/reference_host_kats.json):

* ``group_by_gap``            <- speech_segmentation/backends/ten.py:31-73 (same rule in silero.py:325-361)
* ``pad_and_clamp``           <- speech_segmentation/backends/silero.py:286-297
* ``probs_to_regions``        <- speech_segmentation/backends/whisperseg.py:419-571 (Silero-compatible hysteresis)
* ``vad_looks_broken``        <- modules/vad_failover.py:26-57
* ``LogprobGate``             <- modules/segment_filters.py:80-103 (+ the non-verbal heuristics :105-156)

Plain Python on small lists: this is bookkeeping, not arithmetic; nothing here touches the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

try:  # share the reference's dataclasses when it is importable, so isinstance checks hold inside WhisperJAV
    from whisperjav.modules.speech_segmentation.base import SegmentationResult, SpeechSegment  # type: ignore
except Exception:  # standalone use (this repo's tests, the GPU box)
    @dataclass
    class SpeechSegment:  # speech_segmentation/base.py:13-51
        start_sec: float
        end_sec: float
        start_sample: int = 0
        end_sample: int = 0
        confidence: float = 1.0
        metadata: Dict[str, Any] = field(default_factory=dict)

        @property
        def duration_sec(self) -> float:
            return self.end_sec - self.start_sec

        def to_dict(self) -> Dict[str, Any]:
            return {"start_sec": round(self.start_sec, 3), "end_sec": round(self.end_sec, 3),
                    "duration_sec": round(self.duration_sec, 3), "confidence": round(self.confidence, 3)}

    @dataclass
    class SegmentationResult:  # speech_segmentation/base.py:54-141
        segments: List[SpeechSegment]
        groups: List[List[SpeechSegment]]
        method: str
        audio_duration_sec: float
        parameters: Dict[str, Any]
        processing_time_sec: float = 0.0

        @property
        def speech_coverage_sec(self) -> float:
            return sum(s.duration_sec for s in self.segments)

        @property
        def speech_coverage_ratio(self) -> float:
            return self.speech_coverage_sec / self.audio_duration_sec if self.audio_duration_sec > 0 else 0.0

        @property
        def num_segments(self) -> int:
            return len(self.segments)

        @property
        def num_groups(self) -> int:
            return len(self.groups)

        def to_legacy_format(self) -> List[List[Dict]]:
            return [[{"start": s.start_sample, "end": s.end_sample, "start_sec": s.start_sec, "end_sec": s.end_sec,
                      "metadata": s.metadata} for s in g] for g in self.groups]

        def to_flat_legacy_format(self) -> List[Dict]:
            return [{"start_sec": round(s.start_sec, 3), "end_sec": round(s.end_sec, 3)} for s in self.segments]


def group_by_gap(segments: Sequence[SpeechSegment], max_group_duration_s: float = 29.0,
                 chunk_threshold_s: float = 1.0) -> List[List[SpeechSegment]]:
    """Open a new group when the silence since the previous segment exceeds ``chunk_threshold_s`` or when
    the group would span more than ``max_group_duration_s``."""
    groups: List[List[SpeechSegment]] = []
    prev = None
    for seg in segments:
        if prev is None:
            groups.append([seg])
        else:
            too_far = (seg.start_sec - prev.end_sec) > chunk_threshold_s
            too_long = (seg.end_sec - groups[-1][0].start_sec) > max_group_duration_s
            if too_far or too_long:
                groups.append([seg])
            else:
                groups[-1].append(seg)
        prev = seg
    return groups


def pad_and_clamp(timestamps: Sequence[Dict[str, int]], n_audio: int, start_pad: int = 11200, end_pad: int = 20800) -> List[Tuple[int, int]]:
    """Silero post-VAD sample padding: start - 11200, end + 20800 clamped to ``n_audio - 16``, and a start
    that would land before the previous (already padded) end is pulled up to it."""
    out: List[Tuple[int, int]] = []
    for ts in timestamps:
        a = max(0, int(ts["start"]) - start_pad)
        b = min(n_audio - 16, int(ts["end"]) + end_pad)
        if out and a < out[-1][1]:
            a = out[-1][1]
        out.append((a, b))
    return out


def probs_to_regions(probs: Sequence[float], audio_duration_sec: float, *, frame_ms: float, threshold: float,
                     min_speech_duration_ms: float, min_silence_duration_ms: float, speech_pad_ms: float,
                     max_speech_duration_s: float = 0.0, sample_rate: int = 16000, neg_threshold: Optional[float] = None) -> List[SpeechSegment]:
    """Frame probabilities -> speech regions with Silero-style hysteresis: onset at ``p >= threshold``;
    an offset candidate opens at ``p < neg_threshold`` (default ``max(threshold - 0.15, 0.01)``) and is confirmed once that silence has
    lasted ``min_silence`` frames; regions longer than ``max_speech`` are cut; regions shorter than
    ``min_speech`` are dropped; finally every region is padded by ``speech_pad`` without overlapping its
    neighbours."""
    n = len(probs)
    if n == 0:
        return []
    off_thr = float(neg_threshold) if neg_threshold is not None else max(float(threshold) - 0.15, 0.01)
    min_speech = max(1, int(min_speech_duration_ms / frame_ms))
    min_silence = max(1, int(min_silence_duration_ms / frame_ms))
    pad = max(0, int(speech_pad_ms / frame_ms))
    max_speech = int(max_speech_duration_s * 1000.0 / frame_ms) if max_speech_duration_s and max_speech_duration_s > 0 else n

    regions: List[Dict[str, Any]] = []
    start = -1          # frame where the open region began, -1 when idle
    quiet_since = 0     # first frame of the pending silence (0 = none pending; frame 0 can never be one)
    heard: List[float] = []

    def close(end: int, keep: bool):
        nonlocal start, quiet_since, heard
        if keep:
            vals = heard[: end - start]
            reg = {"start": start, "end": end}
            if vals:
                reg.update(avg=float(sum(vals) / len(vals)), lo=float(min(vals)), hi=float(max(vals)))
            regions.append(reg)
        start, quiet_since, heard = -1, 0, []

    for i in range(n):
        p = float(probs[i])
        if start >= 0:
            heard.append(p)
        if start < 0:
            if p >= threshold:
                start, heard = i, [p]
            continue
        if i - start > max_speech:
            close(start + max_speech, True)
            continue
        if p < off_thr:
            if not quiet_since:
                quiet_since = i
            if i - quiet_since >= min_silence:
                close(quiet_since, quiet_since - start >= min_speech)
        elif p >= threshold and quiet_since:
            quiet_since = 0
    if start >= 0 and n - start >= min_speech:
        regions.append({"start": start, "end": n, "avg": float(sum(heard) / len(heard)), "lo": float(min(heard)), "hi": float(max(heard))})

    for k, reg in enumerate(regions):
        reg["start"] = max(regions[k - 1]["end"] if k else 0, reg["start"] - pad)
        reg["end"] = min(regions[k + 1]["start"] if k + 1 < len(regions) else n, reg["end"] + pad)

    out: List[SpeechSegment] = []
    for reg in regions:
        a = reg["start"] * frame_ms / 1000.0
        b = min(reg["end"] * frame_ms / 1000.0, audio_duration_sec)
        if b <= a:
            continue
        meta = {}
        if "lo" in reg:
            meta = {"min_prob": reg["lo"], "max_prob": reg["hi"]}
        out.append(SpeechSegment(start_sec=a, end_sec=b, start_sample=int(a * sample_rate), end_sample=int(b * sample_rate),
                                 confidence=max(0.0, min(1.0, reg.get("avg", 1.0))), metadata=meta))
    return out


def vad_looks_broken(vad_groups: Optional[Iterable[Iterable[Dict[str, float]]]], audio_duration: float,
                     min_duration_for_fallback: float = 120.0, min_coverage_ratio: float = 0.01) -> bool:
    """True when a long clip (>= 120 s) got no speech, under 1 % coverage, or <= 2 blobs on >= 480 s: the
    caller then transcribes the whole clip instead of trusting the VAD."""
    if audio_duration <= 0 or audio_duration < min_duration_for_fallback:
        return False
    flat = [seg for group in (vad_groups or []) for seg in (group or [])]
    if not flat:
        return True
    speech = 0.0
    for seg in flat:
        a, b = float(seg.get("start_sec", 0.0)), float(seg.get("end_sec", 0.0))
        if b > a:
            speech += b - a
    if speech / audio_duration < min_coverage_ratio:
        return True
    return len(flat) <= 2 and audio_duration >= 4 * min_duration_for_fallback


class LogprobGate:
    """Post-decode segment gate: drop when ``avg_logprob < threshold`` (threshold lowered by ``margin`` for
    segments no longer than ``short_window`` s), optionally drop non-verbal vocalisations."""

    KEYWORDS = ("music", "applause", "laugh", "laughs", "laughter", "sfx", "fx", "noise", "silence", "ambient", "moan", "moans",
                "moaning", "groan", "groans", "sigh", "sighs", "breath", "breathing", "喘", "喘ぎ", "喘ぎ声", "うめき", "うめき声")
    NOTES = set("♪♫")
    VOCAL_CHARS = set("ahmnou" "ぁあァアんンっッふフぅゥうウおオえエはハほホ")
    IGNORED = set("!！?？。、,.・~〜～ー… 　")

    def __init__(self, enabled: bool = True, logprob_threshold: Optional[float] = None, logprob_margin: float = 0.0,
                 drop_nonverbal_vocals: bool = False, short_segment_window: float = 1.6):
        self.enabled = bool(enabled)
        self.threshold = logprob_threshold
        self.margin = max(0.0, logprob_margin or 0.0)
        self.drop_nonverbal = bool(drop_nonverbal_vocals)
        self.short_window = max(0.4, float(short_segment_window or 1.6))

    def should_filter(self, avg_logprob: float, duration: float, text: str):
        if not self.enabled:
            return False, None, None
        eff = self.threshold
        if eff is not None and self.margin > 0 and duration <= self.short_window:
            eff = eff - self.margin
        if eff is not None and avg_logprob < eff:
            return True, "logprob", eff
        if self.drop_nonverbal and self.looks_nonverbal(text):
            return True, "nonverbal", eff
        return False, None, eff

    @classmethod
    def looks_nonverbal(cls, text: str) -> bool:
        t = (text or "").strip()
        if not t:
            return False
        if all(ch in cls.NOTES or ch in cls.IGNORED for ch in t):
            return True
        core = t.lower().strip()
        core = core.lstrip("[](){}<>").rstrip("[](){}<>").strip()
        if not core:
            return False
        if any(k in core for k in cls.KEYWORDS):
            return True
        bare = "".join(ch for ch in core if ch not in cls.IGNORED)
        return bool(bare) and len(bare) <= 6 and all(ch in cls.VOCAL_CHARS for ch in bare)


# ---------------------------------------------------------------------------------------------------------
# TEN-style pipeline: per-hop speech flags -> regions -> merge -> pad -> split -> SpeechSegment
# (speech_segmentation/backends/ten.py:280-515: _flags_to_segments, _merge_by_silence, _apply_padding,
#  _split_long_segments, _even_split, _make_subsegment, _to_speech_segments)
# ---------------------------------------------------------------------------------------------------------
def flags_to_regions(flags: Sequence[int], probs: Sequence[float], frame_s: float, audio_duration: float,
                     min_speech_duration_ms: float = 81, max_speech_duration_s: float = 10.0) -> List[Dict[str, Any]]:
    """Runs of flag == 1 become regions {'start','end','probs'}; a run is cut (and restarted on the same frame) once it
    has lasted ``max_speech_duration_s``; regions shorter than ``min_speech_duration_ms`` are dropped."""
    out: List[Dict[str, Any]] = []
    start = None
    heard: List[float] = []

    def emit(end: float):
        if (end - start) * 1000 >= min_speech_duration_ms:
            out.append({"start": start, "end": end, "probs": list(heard)})

    for i, flag in enumerate(flags):
        t = i * frame_s
        if flag == 1:
            if start is None:
                start, heard = t, [probs[i]]
            else:
                heard.append(probs[i])
                if max_speech_duration_s > 0 and t - start >= max_speech_duration_s:
                    emit(t)
                    start, heard = t, [probs[i]]
        elif start is not None:
            emit(t)
            start, heard = None, []
    if start is not None:
        emit(min(len(flags) * frame_s, audio_duration))
    return out


def merge_close_regions(regions: Sequence[Dict[str, Any]], min_silence_duration_ms: float = 100) -> List[Dict[str, Any]]:
    """Fuse neighbours whose gap is <= ``min_silence_duration_ms`` (probabilities are concatenated)."""
    if not regions or min_silence_duration_ms <= 0:
        return list(regions)
    gap_s = min_silence_duration_ms / 1000.0
    out: List[Dict[str, Any]] = []
    for r in regions:
        if out and r["start"] - out[-1]["end"] <= gap_s:
            out[-1]["end"] = r["end"]
            out[-1]["probs"].extend(r["probs"])
        else:
            out.append({"start": r["start"], "end": r["end"], "probs": list(r["probs"])})
    return out


def pad_regions(regions: Sequence[Dict[str, Any]], audio_duration: float, start_pad_ms: float = 50,
                end_pad_ms: float = 150) -> List[Dict[str, Any]]:
    """Start earlier by ``start_pad_ms``, end later by ``end_pad_ms``; a start never precedes the previous padded end."""
    out: List[Dict[str, Any]] = []
    for i, r in enumerate(regions):
        a = max(0.0, r["start"] - start_pad_ms / 1000.0)
        b = min(audio_duration, r["end"] + end_pad_ms / 1000.0)
        if i > 0 and out and a < out[-1]["end"]:
            a = out[-1]["end"]
        if b > a:
            out.append({"start": a, "end": b, "probs": r["probs"], "raw_start": r["start"], "raw_end": r["end"]})
    return out


def _sub_region(parent: Dict[str, Any], a: float, b: float, frame_s: float) -> Dict[str, Any]:
    i0 = max(0, int((a - parent["start"]) / frame_s)) if frame_s > 0 else 0
    i1 = min(len(parent["probs"]), int((b - parent["start"]) / frame_s)) if frame_s > 0 else 0
    return {"start": a, "end": b, "probs": parent["probs"][i0:i1] if i0 < i1 else [],
            "raw_start": parent.get("raw_start", a), "raw_end": parent.get("raw_end", b)}


def _even_parts(r: Dict[str, Any], max_dur: float) -> List[Dict[str, Any]]:
    import math
    dur = r["end"] - r["start"]
    n = max(1, int(math.ceil(dur / max_dur)))
    part = dur / n
    frame_s = dur / len(r["probs"]) if len(r["probs"]) > 0 else 0.016
    return [_sub_region(r, r["start"] + i * part, min(r["start"] + i * part + part, r["end"]), frame_s) for i in range(n)]


def split_long_regions(regions: Sequence[Dict[str, Any]], max_speech_duration_s: float = 10.0) -> List[SpeechSegment]:
    """Regions longer than ``max_speech_duration_s`` are cut at local minima of the box-smoothed probability curve
    (a cut only once 80 % of the limit has elapsed since the last one), else evenly."""
    import numpy as np
    final: List[Dict[str, Any]] = []
    for r in regions:
        dur = r["end"] - r["start"]
        if max_speech_duration_s <= 0 or dur <= max_speech_duration_s:
            final.append(r)
            continue
        p = r["probs"]
        if len(p) < 2:
            final.extend(_even_parts(r, max_speech_duration_s))
            continue
        arr = np.array(p, dtype=np.float32)
        win = max(3, len(arr) // 20)
        smooth = np.convolve(arr, np.ones(win) / win, mode="same")
        minima = [j for j in range(1, len(smooth) - 1) if smooth[j] <= smooth[j - 1] and smooth[j] <= smooth[j + 1]]
        frame_s = dur / len(p)
        cuts, last = [], r["start"]
        for j in minima:
            t = r["start"] + j * frame_s
            if t - last > max_speech_duration_s * 0.8:
                cuts.append(t)
                last = t
        if not minima or not cuts:
            final.extend(_even_parts(r, max_speech_duration_s))
            continue
        prev = r["start"]
        for t in cuts:
            if t - prev > 0.05:
                final.append(_sub_region(r, prev, t, frame_s))
                prev = t
        if r["end"] - prev > 0.05:
            final.append(_sub_region(r, prev, r["end"], frame_s))
    out: List[SpeechSegment] = []
    for r in final:
        p = r.get("probs", [])
        out.append(SpeechSegment(start_sec=r["start"], end_sec=r["end"], start_sample=int(r["start"] * 16000),
                                 end_sample=int(r["end"] * 16000), confidence=(sum(p) / len(p)) if p else 1.0,
                                 metadata={"raw_start": r.get("raw_start", r["start"]), "raw_end": r.get("raw_end", r["end"])}))
    return out


def ten_style_segments(flags: Sequence[int], probs: Sequence[float], audio_duration: float, *, hop_size: int = 256,
                       sample_rate: int = 16000, min_speech_duration_ms: float = 81, min_silence_duration_ms: float = 100,
                       max_speech_duration_s: float = 10.0, start_pad_ms: float = 50, end_pad_ms: float = 150) -> List[SpeechSegment]:
    """detect -> merge -> pad -> split, exactly the order of TenSpeechSegmenter.segment (ten.py:241-251)."""
    frame_s = hop_size / sample_rate
    r = flags_to_regions(flags, probs, frame_s, audio_duration, min_speech_duration_ms, max_speech_duration_s)
    r = merge_close_regions(r, min_silence_duration_ms)
    r = pad_regions(r, audio_duration, start_pad_ms, end_pad_ms)
    return split_long_regions(r, max_speech_duration_s)


# ----------------------------------------------------------------------------- beam search, host half
def beam_finalize_and_rank(finished: Sequence[Tuple[Sequence[int], float]], live: Sequence[Tuple[Sequence[int], float]], beam_size: int,
                           n_initial: int, eot: int, length_penalty: Optional[float] = None) -> Tuple[List[int], float]:
    """openai-whisper decoding.py::BeamSearchDecoder.finalize + MaximumLikelihoodRanker.rank for one window.

    ``finished``: (token sequence incl. the initial tokens and the closing EOT, sum_logprob) in the order the device collected
    them; ``live``: the beams still running (token sequence without EOT, sum_logprob).  Fewer than ``beam_size`` finished
    sequences are topped up with the best live beams, EOT appended.  Returns the sampled tokens (initial tokens and everything
    from the first EOT on stripped) of the candidate with the best ``sum_logprob / length`` (or the Google-NMT penalty
    ``((5 + length) / 6) ** length_penalty``) and its sum_logprob."""
    seqs: Dict[Tuple[int, ...], float] = {}
    for seq, score in finished:
        seqs[tuple(int(t) for t in seq)] = float(score)
    if len(seqs) < beam_size:
        order = sorted(range(len(live)), key=lambda j: live[j][1])[::-1]  # np.argsort(sum_logprobs)[::-1]
        for j in order:
            seqs[tuple(int(t) for t in live[j][0]) + (eot,)] = float(live[j][1])
            if len(seqs) >= beam_size:
                break
    best, best_norm, best_score = None, None, 0.0
    for seq, score in seqs.items():
        body = list(seq[n_initial:])
        if eot in body:
            body = body[: body.index(eot)]
        length = len(body)
        penalty = length if length_penalty is None else ((5 + length) / 6) ** length_penalty
        norm = score / penalty if penalty else float("-inf")
        if best_norm is None or norm > best_norm:  # np.argmax: the first maximum wins
            best, best_norm, best_score = body, norm, score
    return best if best is not None else [], best_score
