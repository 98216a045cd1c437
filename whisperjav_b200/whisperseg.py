"""WhisperSeg-class speech segmenter on the GPU: the drop-in for ``WhisperSegSpeechSegmenter``
(whisperjav/modules/speech_segmentation/backends/whisperseg.py), the reference's default segmenter on the ensemble / qwen /
decoupled paths (whisper_pro_asr.py:66).

Reference pipeline (whisperseg.py:355-413, 419-571, 577-628) and where each step runs here:

  1. 30 s chunks, zero-padded                     -> host slicing into one batch
  2. 80-bin log-mel via WhisperFeatureExtractor   -> ``wjb_logmel_f16`` with HF semantics (reflect at 480000; pinned to HF's own
                                                     extractor by tests/golden/hf_logmel_80.npz) -- the second consumer of the
                                                     mel kernel
  3. ONNX encoder-decoder -> 1500 frame logits    -> Whisper-base-shaped encoder (``wjb_encoder_forward``, tcgen05 GEMMs + flash
                                                     attention) + a frame head (``wjb_frame_head_f16``: Linear(512 -> 1), sigmoid)
  4. sigmoid, Silero-compatible state machine     -> ``hostlogic.probs_to_regions`` (pinned by KATs generated from the
                                                     reference's own ``_probs_to_segments``)
  5. ``group_segments``                           -> ``hostlogic.group_by_gap`` (KATs from the reference)

All chunks of all clips of a call go through the device as ONE batch (the reference runs one ONNX call per chunk).

The vendor network (TransWithAI/Whisper-Vad-EncDec-ASMR-onnx) is not available offline and its decoder is not described in the
reference; the network here is Whisper-base's encoder with a per-frame linear head, seeded random-init unless a ``state_dict``
(openai / HF encoder names + ``head.weight`` [512], ``head.bias`` []) is given.  Numerics are therefore UNPINNED (as the
reference's own tests leave them); the boundary, the feature path and the post-processing are pinned.
"""
from __future__ import annotations

import threading
import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from . import hostlogic as H
from .audioio import read_wav_mono

SR = 16000
CHUNK = 480000
FRAMES = 1500
FRAME_MS = 20.0


class B200WhisperSegSegmenter:
    """Same constructor keywords and defaults as the reference's WhisperSegSpeechSegmenter (whisperseg.py:80-141)."""

    def __init__(self, threshold: float = 0.35, min_speech_duration_ms: int = 100, min_silence_duration_ms: int = 100,
                 speech_pad_ms: int = 300, max_speech_duration_s: Optional[float] = None, chunk_threshold_s: Optional[float] = 1.0,
                 max_group_duration_s: Optional[float] = None, device: str = "cuda", state_dict: Optional[dict] = None, seed: int = 21,
                 max_batch: int = 64, **kwargs: Any):
        self.threshold = float(threshold)
        self.min_speech_duration_ms = int(min_speech_duration_ms)
        self.min_silence_duration_ms = int(min_silence_duration_ms)
        self.speech_pad_ms = int(speech_pad_ms)
        if chunk_threshold_s is not None:
            self.chunk_threshold_s = float(chunk_threshold_s)
        elif "chunk_threshold" in kwargs:
            self.chunk_threshold_s = float(kwargs["chunk_threshold"])
        else:
            self.chunk_threshold_s = 1.0
        self.max_group_duration_s = float(max_group_duration_s) if max_group_duration_s is not None else 29.0
        self.max_speech_duration_s = float(max_speech_duration_s) if max_speech_duration_s is not None else self.max_group_duration_s
        self._device, self._sd, self._seed, self._max_batch = device, state_dict, int(seed), int(max_batch)
        self._model = None
        self._head = None
        self._lock = threading.Lock()

    @property
    def name(self) -> str:
        return "b200-whisperseg"

    @property
    def display_name(self) -> str:
        return "B200 WhisperSeg-class VAD (Whisper-base encoder + frame head, CUDA)"

    def get_supported_sample_rates(self) -> List[int]:
        return [SR]

    def cleanup(self) -> None:
        with self._lock:
            if self._model is not None:
                self._model.close()
            self._model = None
            self._head = None

    def _get_parameters(self) -> Dict[str, Any]:
        return {"threshold": self.threshold, "min_speech_duration_ms": self.min_speech_duration_ms,
                "min_silence_duration_ms": self.min_silence_duration_ms, "speech_pad_ms": self.speech_pad_ms,
                "max_speech_duration_s": self.max_speech_duration_s, "chunk_threshold_s": self.chunk_threshold_s,
                "max_group_duration_s": self.max_group_duration_s, "device": "GPU (CUDA, libwjb200)", "frame_duration_ms": int(FRAME_MS),
                "chunk_duration_ms": 30000}

    def _ensure_model(self):
        if self._model is None:
            with self._lock:
                if self._model is None:
                    import torch
                    from .model import WhisperB200
                    from .synth import DIMS, synth_weights
                    dims = DIMS["base"]
                    sd = dict(self._sd) if self._sd is not None else synth_weights(dims, seed=self._seed)
                    g = torch.Generator().manual_seed(self._seed + 1)
                    hw = sd.pop("head.weight", None)
                    hb = sd.pop("head.bias", None)
                    if hw is None:  # seeded stand-in: a head whose logit spread straddles the default threshold
                        hw = torch.randn(dims.n_audio_state, generator=g) * (2.0 / dims.n_audio_state ** 0.5)
                        hb = torch.tensor(0.0)
                    if self._sd is not None and not any(k.startswith("decoder.") or k.startswith("model.decoder.") for k in sd):
                        sd.update({k: v for k, v in synth_weights(dims, seed=self._seed).items() if k.startswith("decoder.")})  # unused filler
                    m = WhisperB200(dims, sd, device=self._device, max_batch=self._max_batch)
                    self._head = (hw.reshape(-1).to(torch.float16).to(m.device), float(hb))
                    self._model = m
        return self._model

    # ------------------------------------------------------------------ device stage
    def frame_probs(self, clips: Sequence[np.ndarray]) -> List[np.ndarray]:
        """Per-clip frame probabilities (20 ms frames, 1500 per 30 s chunk), every chunk of every clip in shared device batches."""
        import torch
        from . import _lib
        m = self._ensure_model()
        index = []  # (clip, chunk)
        for ci, a in enumerate(clips):
            for k in range(0, len(a), CHUNK):
                index.append((ci, k))
        out: List[List[np.ndarray]] = [[] for _ in clips]
        hw, hb = self._head
        for c0 in range(0, len(index), self._max_batch):
            part = index[c0: c0 + self._max_batch]
            host = torch.zeros(len(part), CHUNK, dtype=torch.float32)
            for j, (ci, k) in enumerate(part):
                seg = clips[ci][k: k + CHUNK]
                host[j, : len(seg)] = torch.from_numpy(np.ascontiguousarray(seg, dtype=np.float32))
            audio = host.to(m.device)
            ns = torch.full((len(part),), CHUNK, dtype=torch.int32, device=m.device)  # HF: the zero padding is part of the signal
            mel = m.log_mel(audio, ns, n_frames=3000, layout="time", reflect_total=CHUNK)
            xa = m.encode(mel)
            probs = torch.empty(len(part), FRAMES, dtype=torch.float32, device=m.device)
            with torch.cuda.device(m.device):
                _lib.check(m.lib.wjb_frame_head_f16(_lib.ptr(xa), _lib.ptr(hw), hb, _lib.ptr(probs), len(part) * FRAMES, m.dims.n_audio_state,
                                                   _lib.stream_ptr()), "wjb_frame_head_f16")
            h = probs.cpu().numpy()
            for j, (ci, _) in enumerate(part):
                out[ci].append(h[j])
        return [np.concatenate(p) if p else np.zeros(0, np.float32) for p in out]

    # ------------------------------------------------------------------ public surface
    def segment(self, audio: Union[np.ndarray, Path, str], sample_rate: int = SR, **kwargs: Any):
        return self.segment_batch([audio], sample_rate=sample_rate, **kwargs)[0]

    def segment_batch(self, audios: Sequence[Union[np.ndarray, Path, str]], sample_rate: int = SR, **kwargs: Any):
        t0 = time.time()
        clips, durations = [], []
        for a in audios:
            if isinstance(a, (str, Path)):
                data, sr = read_wav_mono(a)
            else:
                data, sr = np.asarray(a), sample_rate
            if data.ndim > 1:
                data = data.mean(axis=0 if data.shape[0] > data.shape[1] else 1)
            data = data.astype(np.float32, copy=False)
            durations.append(len(data) / sr if sr > 0 else 0.0)
            if sr != SR and len(data):  # whisperseg.py:674-689 (scipy.signal.resample)
                from scipy import signal
                data = signal.resample(data, int(len(data) * SR / sr)).astype(np.float32)
            clips.append(data)
        try:
            probs = self.frame_probs(clips)
        except Exception:  # the reference returns an empty result when inference fails (whisperseg.py:603-614)
            import logging
            logging.getLogger("whisperjav_b200").error("b200-whisperseg inference failed", exc_info=True)
            return [H.SegmentationResult([], [], self.name, d, self._get_parameters(), time.time() - t0) for d in durations]
        return [self._postprocess(p, d, time.time() - t0) for p, d in zip(probs, durations)]

    def _postprocess(self, probs: np.ndarray, duration: float, elapsed: float):
        segs = H.probs_to_regions(probs, duration, frame_ms=FRAME_MS, threshold=self.threshold,
                                  min_speech_duration_ms=self.min_speech_duration_ms, min_silence_duration_ms=self.min_silence_duration_ms,
                                  speech_pad_ms=self.speech_pad_ms, max_speech_duration_s=self.max_speech_duration_s)
        groups = H.group_by_gap(segs, self.max_group_duration_s, self.chunk_threshold_s)
        return H.SegmentationResult(segs, groups, self.name, duration, self._get_parameters(), elapsed)
