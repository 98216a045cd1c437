"""TEST INFRASTRUCTURE -- CPU restatement of the energy tokenizer behind the reference's scene detector.

The reference calls ``auditok.split(audio_bytes, sampling_rate=…, channels=1, sample_width=2, min_dur=…, max_dur=…,
max_silence=…, energy_threshold=…, drop_trailing_silence=True)`` twice per film
(whisperjav/modules/scene_detection_backends/auditok_backend.py:379-392 and :552-567).  ``auditok`` is a third-party
dependency (pinned ``auditok==0.3.0`` in the reference's install_log.txt:311) that is ABSENT from /root/reference and from
this container, so this file restates its published algorithm: ``auditok.core.split`` -> ``AudioReader`` blocks of
``analysis_window`` = 0.05 s -> ``AudioEnergyValidator`` (``signal.calculate_energy``: 20 log10 of the float64 RMS of the
int16 samples, clipped at 1e-10) -> ``StreamTokenizer`` (states SILENCE / POSSIBLE_SILENCE / POSSIBLE_NOISE / NOISE,
``DROP_TRAILING_SILENCE`` mode) -> ``AudioRegion`` start/end.

PARITY UNPINNED at the auditok boundary (no auditok here, no golden vectors in the reference's tests).  What IS pinned:
the reference's own two-pass driver (``AuditokSceneDetector.detect_scenes``) is executed here with this module standing
in for ``auditok`` (tests/golden/make_scene_kats.py) and the product must reproduce its scenes exactly.

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

DEFAULT_ANALYSIS_WINDOW = 0.05
_EPSILON = 1e-10


def calculate_energy(x: np.ndarray) -> float:
    """auditok/signal.py::calculate_energy for one channel: x int16 samples of one block."""
    xf = np.asarray(x).astype("float64")
    energy_sqrt = math.sqrt(float(np.mean(xf ** 2))) if xf.size else 0.0
    energy_sqrt = max(energy_sqrt, _EPSILON)
    return 20.0 * math.log10(energy_sqrt)


def duration_to_nb_windows(duration: float, analysis_window: float, round_fn=round, epsilon: float = 0.0) -> int:
    """auditok/core.py::_duration_to_nb_windows"""
    if duration < 0 or analysis_window <= 0:
        raise ValueError("'duration' (%r) must be >= 0 and 'analysis_window' (%r) > 0" % (duration, analysis_window))
    if duration == 0:
        return 0
    return int(round_fn(duration / analysis_window + epsilon))


class StreamTokenizer:
    """auditok/core.py::StreamTokenizer, frame lists and all (the product keeps counters only)."""
    SILENCE, POSSIBLE_SILENCE, POSSIBLE_NOISE, NOISE = 0, 1, 2, 3
    NORMAL, STRICT_MIN_LENGTH, DROP_TRAILING_SILENCE = 0, 2, 4

    def __init__(self, is_valid, min_length, max_length, max_continuous_silence, init_min=0, init_max_silence=0, mode=0):
        if max_length <= 0:
            raise ValueError("'max_length' must be > 0")
        if min_length <= 0 or min_length > max_length:
            raise ValueError("'min_length' must be > 0 and <= 'max_length'")
        if max_continuous_silence >= max_length:
            raise ValueError("'max_continuous_silence' must be < 'max_length'")
        if init_min >= max_length:
            raise ValueError("'init_min' must be < 'max_length'")
        self.is_valid = is_valid
        self.min_length, self.max_length = min_length, max_length
        self.max_continuous_silence = max_continuous_silence
        self.init_min, self.init_max_silent = init_min, init_max_silence
        self._strict_min_length = (mode & self.STRICT_MIN_LENGTH) != 0
        self._drop_trailing_silence = (mode & self.DROP_TRAILING_SILENCE) != 0
        self._reinitialize()

    def _reinitialize(self):
        self._contiguous_token = False
        self._data = []
        self._tokens = []
        self._state = self.SILENCE
        self._current_frame = -1
        self._init_count = 0
        self._silence_length = 0
        self._start_frame = 0

    def tokenize(self, frames):
        self._reinitialize()
        out = []
        for frame in frames:
            self._current_frame += 1
            token = self._process(frame)
            if token is not None:
                out.append(token)
        token = self._post_process()
        if token is not None:
            out.append(token)
        return out

    def _process(self, frame):
        frame_is_valid = self.is_valid(frame)
        if self._state == self.SILENCE:
            if frame_is_valid:
                self._init_count = 1
                self._silence_length = 0
                self._start_frame = self._current_frame
                self._data.append(frame)
                if self._init_count >= self.init_min:
                    self._state = self.NOISE
                    if len(self._data) >= self.max_length:
                        return self._process_end_of_detection(True)
                else:
                    self._state = self.POSSIBLE_NOISE
        elif self._state == self.POSSIBLE_NOISE:
            if frame_is_valid:
                self._silence_length = 0
                self._init_count += 1
                self._data.append(frame)
                if self._init_count >= self.init_min:
                    self._state = self.NOISE
                    if len(self._data) >= self.max_length:
                        return self._process_end_of_detection(True)
            else:
                self._silence_length += 1
                if self._silence_length > self.init_max_silent or len(self._data) + 1 >= self.max_length:
                    self._data = []
                    self._state = self.SILENCE
                else:
                    self._data.append(frame)
        elif self._state == self.NOISE:
            if frame_is_valid:
                self._data.append(frame)
                if len(self._data) >= self.max_length:
                    return self._process_end_of_detection(True)
            elif self.max_continuous_silence <= 0:
                self._state = self.SILENCE
                return self._process_end_of_detection()
            else:
                self._silence_length = 1
                self._data.append(frame)
                self._state = self.POSSIBLE_SILENCE
                if len(self._data) == self.max_length:
                    return self._process_end_of_detection(True)
        elif self._state == self.POSSIBLE_SILENCE:
            if frame_is_valid:
                self._data.append(frame)
                self._silence_length = 0
                self._state = self.NOISE
                if len(self._data) >= self.max_length:
                    return self._process_end_of_detection(True)
            else:
                if self._silence_length >= self.max_continuous_silence:
                    self._state = self.SILENCE
                    if self._silence_length < len(self._data):
                        return self._process_end_of_detection()
                    self._data = []
                    self._silence_length = 0
                else:
                    self._data.append(frame)
                    self._silence_length += 1
                    if len(self._data) >= self.max_length:
                        return self._process_end_of_detection(True)
        return None

    def _post_process(self):
        if self._state in (self.NOISE, self.POSSIBLE_SILENCE):
            if len(self._data) > 0 and len(self._data) > self._silence_length:
                return self._process_end_of_detection()
        return None

    def _process_end_of_detection(self, truncated=False):
        if not truncated and self._drop_trailing_silence and self._silence_length > 0:
            self._data = self._data[0: -self._silence_length]
        if (len(self._data) >= self.min_length) or (len(self._data) > 0 and not self._strict_min_length and self._contiguous_token):
            start_frame = self._start_frame
            end_frame = self._start_frame + len(self._data) - 1
            data = self._data
            self._data = []
            if truncated:
                self._start_frame = self._current_frame + 1
                self._contiguous_token = True
            else:
                self._contiguous_token = False
            return data, start_frame, end_frame
        self._contiguous_token = False
        self._data = []
        return None


class Region:
    """The two attributes of ``auditok.AudioRegion`` the reference reads (auditok_backend.py:262-264,425-427)."""

    def __init__(self, start: float, end: float):
        self.start, self.end = start, end

    def __repr__(self):
        return "Region(%.3f, %.3f)" % (self.start, self.end)


def split(input, min_dur=0.2, max_dur=5, max_silence=0.3, drop_trailing_silence=False, strict_min_dur=False, **kwargs) -> List[Region]:
    """auditok/core.py::split for raw PCM16 mono bytes (the only form the reference passes)."""
    if min_dur <= 0:
        raise ValueError("'min_dur' must be > 0")
    if max_dur <= 0:
        raise ValueError("'max_dur' must be > 0")
    if max_silence < 0:
        raise ValueError("'max_silence' must be >= 0")
    sr = int(kwargs["sampling_rate"])
    assert int(kwargs.get("channels", 1)) == 1 and int(kwargs.get("sample_width", 2)) == 2
    analysis_window = kwargs.get("analysis_window", DEFAULT_ANALYSIS_WINDOW)
    energy_threshold = kwargs.get("energy_threshold", 50)
    samples = np.frombuffer(input, dtype="<i2")
    block_size = int(analysis_window * sr)
    if block_size == 0:
        raise ValueError("too small analysis window")
    block_dur = block_size / sr
    mode = StreamTokenizer.DROP_TRAILING_SILENCE if drop_trailing_silence else 0
    if strict_min_dur:
        mode |= StreamTokenizer.STRICT_MIN_LENGTH
    min_length = duration_to_nb_windows(min_dur, analysis_window, math.ceil)
    max_length = duration_to_nb_windows(max_dur, analysis_window, math.floor, _EPSILON)
    max_continuous_silence = duration_to_nb_windows(max_silence, analysis_window, math.floor, _EPSILON)
    if min_length > max_length:
        raise ValueError("'min_dur' (%r) results in more analysis windows than 'max_dur' (%r)" % (min_dur, max_dur))
    if max_continuous_silence >= max_length:
        raise ValueError("'max_silence' (%r) must be < 'max_dur' (%r) in analysis windows" % (max_silence, max_dur))
    frames = [samples[i: i + block_size] for i in range(0, len(samples), block_size)]
    tok = StreamTokenizer(lambda fr: calculate_energy(fr) >= energy_threshold, min_length, max_length, max_continuous_silence, mode=mode)
    regions = []
    for data, start_frame, _ in tok.tokenize(frames):
        start = start_frame * block_dur
        n = sum(len(d) for d in data)
        regions.append(Region(start, start + n / sr))
    return regions


# ------------------------------------------------------------------------------------------------
# The reference's two-pass driver (auditok_backend.py:229-524), restated for arrays in memory.
# ------------------------------------------------------------------------------------------------

def brute_force_split(start_sec, end_sec, chunk_duration, min_duration=0.3):
    """scene_detection_backends/utils.py:155-200"""
    out = []
    total = end_sec - start_sec
    if total <= 0:
        return out
    n = int(np.ceil(total / max(chunk_duration, min_duration)))
    for i in range(n):
        a = start_sec + i * chunk_duration
        b = min(start_sec + (i + 1) * chunk_duration, end_sec)
        if b - a < min_duration:
            continue
        out.append((a, b))
    return out


def detect_scenes(audio: np.ndarray, sr: int, *, max_duration=29.0, min_duration=0.2, pass1_min_duration=0.3, pass1_max_duration=2700.0,
                  pass1_max_silence=1.8, pass1_energy_threshold=32, pass2_min_duration=0.3, pass2_max_duration: Optional[float] = None,
                  pass2_max_silence=0.94, pass2_energy_threshold=38, brute_force_fallback=True, brute_force_chunk_s: Optional[float] = None,
                  pad_edges_s=0.0):
    """-> (scenes [(start, end, pass, method)], story_lines [(start, end)])"""
    if pass2_max_duration is None:
        pass2_max_duration = max(max_duration - 1.0, min_duration)
    if brute_force_chunk_s is None:
        brute_force_chunk_s = max_duration
    total = len(audio) / sr

    def clamp(s, e):
        s2 = max(0.0, s - pad_edges_s)
        e2 = min(total, e + pad_edges_s)
        return s2, max(e2, s2)

    audio = np.asarray(audio, dtype=np.float32)
    story = split((audio * 32767).astype(np.int16).tobytes(), sampling_rate=sr, channels=1, sample_width=2, min_dur=pass1_min_duration,
                  max_dur=pass1_max_duration, max_silence=min(total * 0.95, pass1_max_silence), energy_threshold=pass1_energy_threshold,
                  drop_trailing_silence=True)
    scenes = []
    for region in story:
        rs, re_ = region.start, region.end
        dur = re_ - rs
        if min_duration <= dur <= max_duration:
            s, e = clamp(rs, re_)
            scenes.append((s, e, 1, "direct"))
            continue
        a, b = int(rs * sr), int(re_ * sr)
        subs = split((audio[a:b] * 32767).astype(np.int16).tobytes(), sampling_rate=sr, channels=1, sample_width=2, min_dur=pass2_min_duration,
                     max_dur=pass2_max_duration, max_silence=min(dur * 0.95, pass2_max_silence), energy_threshold=pass2_energy_threshold,
                     drop_trailing_silence=True)
        if subs:
            for sub in subs:
                ss, se = rs + sub.start, rs + sub.end
                if se - ss < min_duration:
                    continue
                s, e = clamp(ss, se)
                scenes.append((s, e, 2, "granular"))
        elif brute_force_fallback:
            for (ba, bb) in brute_force_split(rs, re_, brute_force_chunk_s, min_duration):
                s, e = clamp(ba, bb)
                scenes.append((s, e, 2, "brute_force"))
    return scenes, [(r.start, r.end) for r in story]


def window_sumsq(audio: np.ndarray, regions: Sequence[tuple], window: int) -> List[np.ndarray]:
    """CPU twin of csrc/scene.cu for the tests: exact per-window sum of squares of the int16-converted samples."""
    out = []
    a = np.asarray(audio, dtype=np.float32)
    for (start, length) in regions:
        q = (a[start: start + length] * 32767).astype(np.int16).astype(np.int64)
        n = (len(q) + window - 1) // window
        out.append(np.array([int(np.sum(q[i * window: (i + 1) * window] ** 2)) for i in range(n)], dtype=np.uint64))
    return out
