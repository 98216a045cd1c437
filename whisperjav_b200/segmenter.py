"""B4 -- SpeechSegmenter backend "b200-vad": the reference's speech-segmentation plug-in surface
(whisperjav/modules/speech_segmentation/base.py:144-203) on top of the GPU VAD kernel.

``segment()`` keeps the Silero backend's contract (backends/silero.py:214-323): per-window probabilities
-> hysteresis state machine -> sample padding / clamp -> gap/duration grouping, returning a
``SegmentationResult``; inference failures raise (Silero semantics, whisper_pro_asr.py:365-367).
``segment_batch`` is the extension that lets many scenes share one device pass.
"""
from __future__ import annotations

import threading
import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from . import hostlogic as H
from .audioio import read_wav_mono

VAD_SR = 16000
WINDOW = 512


class B200SpeechSegmenter:
    """Ctor absorbs ``**kwargs`` (the factory injects version/variant, factory.py:466-483)."""

    def __init__(self, threshold: float = 0.5, min_speech_duration_ms: int = 150, min_silence_duration_ms: int = 300,
                 speech_pad_ms: int = 30, chunk_threshold_s: Optional[float] = None, max_group_duration_s: Optional[float] = None,
                 max_speech_duration_s: Optional[float] = None, start_pad_samples: int = 11200, end_pad_samples: int = 20800,
                 device: str = "cuda", vad_state_dict: Optional[dict] = None, style: str = "silero",
                 start_pad_ms: int = 50, end_pad_ms: int = 150, **kwargs: Any):
        self.threshold = float(threshold)
        self.min_speech_duration_ms = int(min_speech_duration_ms)
        self.min_silence_duration_ms = int(min_silence_duration_ms)
        self.speech_pad_ms = int(speech_pad_ms)
        if chunk_threshold_s is None:
            chunk_threshold_s = kwargs.get("chunk_threshold", 4.0)  # legacy alias (silero.py:163-170)
        self.chunk_threshold_s = float(chunk_threshold_s)
        self.max_group_duration_s = float(max_group_duration_s) if max_group_duration_s is not None else 29.0
        self.max_speech_duration_s = float(max_speech_duration_s) if max_speech_duration_s not in (None, float("inf")) else 0.0
        self.start_pad_samples = int(start_pad_samples)
        self.end_pad_samples = int(end_pad_samples)
        if style not in ("silero", "ten"):
            raise ValueError("style must be 'silero' (hysteresis + sample padding) or 'ten' (flag runs + merge/pad/split)")
        self.style = style
        self.start_pad_ms, self.end_pad_ms = int(start_pad_ms), int(end_pad_ms)
        self._device = device
        self._sd = vad_state_dict
        self._model = None
        self._lock = threading.Lock()

    @property
    def name(self) -> str:
        return "b200-vad"

    @property
    def display_name(self) -> str:
        return "B200 VAD (Silero-class, CUDA)"

    def get_supported_sample_rates(self) -> List[int]:
        return [VAD_SR]

    def _ensure_model(self):
        if self._model is None:
            with self._lock:
                if self._model is None:
                    from .vad import VadB200
                    self._model = VadB200(self._sd, device=self._device)
        return self._model

    def cleanup(self) -> None:
        with self._lock:
            self._model = None

    def _get_parameters(self) -> Dict[str, Any]:
        return {"threshold": self.threshold, "min_speech_duration_ms": self.min_speech_duration_ms,
                "min_silence_duration_ms": self.min_silence_duration_ms, "speech_pad_ms": self.speech_pad_ms,
                "chunk_threshold_s": self.chunk_threshold_s, "max_group_duration_s": self.max_group_duration_s,
                "start_pad_samples": self.start_pad_samples, "end_pad_samples": self.end_pad_samples}

    # ------------------------------------------------------------------ public surface
    def segment(self, audio: Union[np.ndarray, Path, str], sample_rate: int = 16000, **kwargs: Any):
        return self.segment_batch([audio], sample_rate=sample_rate, **kwargs)[0]

    def segment_batch(self, audios: Sequence[Union[np.ndarray, Path, str]], sample_rate: int = 16000, **kwargs: Any):
        import torch
        t0 = time.time()
        model = self._ensure_model()
        clips, durations = [], []
        for a in audios:
            if isinstance(a, (str, Path)):
                data, sr = read_wav_mono(a)
            else:
                data, sr = np.asarray(a, dtype=np.float32), sample_rate
            durations.append(len(data) / sr if sr else 0.0)
            if sr != VAD_SR:  # nearest-index decimation, as backends/silero.py:411-414
                idx = np.linspace(0, len(data) - 1, int(len(data) * VAD_SR / sr)).astype(int)
                data = data[idx]
            clips.append(np.ascontiguousarray(data, dtype=np.float32))
        n = len(clips)
        S = max(max((len(c) for c in clips), default=0), WINDOW)
        host = torch.zeros(n, S, dtype=torch.float32)
        for i, c in enumerate(clips):
            host[i, : len(c)] = torch.from_numpy(c)
        ns = torch.tensor([len(c) for c in clips], dtype=torch.int32)
        probs = model.probs(host.to(model.device), ns.to(model.device)).cpu().numpy()
        out = []
        for i, c in enumerate(clips):
            out.append(self._postprocess(probs[i, : (len(c) + WINDOW - 1) // WINDOW], len(c), durations[i], kwargs, time.time() - t0))
        return out

    def _postprocess(self, probs: np.ndarray, n_audio: int, duration: float, kw: Dict[str, Any], elapsed: float):
        if self.style == "ten":  # backends/ten.py pipeline on 512-sample hops: flags = prob >= threshold
            thr = kw.get("threshold", self.threshold)
            flags = [1 if float(p) >= thr else 0 for p in probs]
            segs = H.ten_style_segments(flags, [float(p) for p in probs], n_audio / VAD_SR, hop_size=WINDOW,
                                        min_speech_duration_ms=kw.get("min_speech_duration_ms", self.min_speech_duration_ms),
                                        min_silence_duration_ms=kw.get("min_silence_duration_ms", self.min_silence_duration_ms),
                                        max_speech_duration_s=self.max_speech_duration_s or 10.0,
                                        start_pad_ms=self.start_pad_ms, end_pad_ms=self.end_pad_ms)
            groups = H.group_by_gap(segs, self.max_group_duration_s, self.chunk_threshold_s)
            return H.SegmentationResult(segs, groups, self.name, duration, self._get_parameters(), elapsed)
        regions = H.probs_to_regions(
            probs, n_audio / VAD_SR, frame_ms=1000.0 * WINDOW / VAD_SR, threshold=kw.get("threshold", self.threshold),
            min_speech_duration_ms=kw.get("min_speech_duration_ms", self.min_speech_duration_ms),
            min_silence_duration_ms=kw.get("min_silence_duration_ms", self.min_silence_duration_ms),
            speech_pad_ms=kw.get("speech_pad_ms", self.speech_pad_ms), max_speech_duration_s=self.max_speech_duration_s)
        if not regions:
            return H.SegmentationResult([], [], self.name, duration, self._get_parameters(), elapsed)
        stamps = [{"start": r.start_sample, "end": r.end_sample} for r in regions]
        padded = H.pad_and_clamp(stamps, n_audio, self.start_pad_samples, self.end_pad_samples)
        segs = [H.SpeechSegment(start_sec=a / VAD_SR, end_sec=b / VAD_SR, start_sample=a, end_sample=b, confidence=r.confidence,
                                metadata=dict(r.metadata)) for (a, b), r in zip(padded, regions) if b > a]
        groups = H.group_by_gap(segs, self.max_group_duration_s, self.chunk_threshold_s)
        return H.SegmentationResult(segs, groups, self.name, duration, self._get_parameters(), elapsed)
