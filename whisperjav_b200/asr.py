"""B1 -- ASR wrapper with the duck-type the fidelity / balanced pipelines construct
(``WhisperProASR``, whisperjav/modules/whisper_pro_asr.py:29-576): same constructor
``(model_config, params, task, tracer=None)``, ``transcribe`` / ``transcribe_to_srt`` / statistics /
``cleanup`` methods and result dicts, but all VAD groups of a scene go through the model as **one device
batch** (the reference calls ``whisper_model.transcribe`` once per group, :306-314)."""
from __future__ import annotations

import gc
import json
import logging
from pathlib import Path
from typing import Dict, List, Optional, Union

import numpy as np

from . import hostlogic as H
from .audioio import compose_srt, read_wav_mono

logger = logging.getLogger("whisperjav")


class _NullTracer:  # utils/parameter_tracer.py NullTracer stand-in
    def emit_transcribe_params(self, **kwargs):
        pass

    def emit(self, *a, **k):
        pass


class B200WhisperASR:
    """Drop-in for ``WhisperProASR`` backed by libwjb200.so."""

    def __init__(self, model_config: Dict, params: Dict, task: str, tracer=None):
        self.tracer = tracer if tracer is not None else _NullTracer()
        self.model_name = model_config.get("model_name", "large-v2")
        self.device = model_config.get("device") or "cuda"
        if self.device == "auto":
            self.device = "cuda"
        decoder_params = dict(params["decoder"])
        vad_params = dict(params["vad"])
        provider_params = dict(params["provider"])
        seg_cfg = dict(params.get("speech_segmenter", {}))
        backend = seg_cfg.get("backend", "b200-vad")
        # constructor firewall (whisper_pro_asr.py:75-94): resolver-produced Silero presets only reach Silero-like backends
        if not (backend.startswith("silero") or backend == "b200-vad"):
            vad_params = {}
        self.vad_threshold = vad_params.get("threshold", 0.4)
        self.min_speech_duration_ms = vad_params.get("min_speech_duration_ms", 150)
        self.vad_chunk_threshold = vad_params.get("chunk_threshold", 4.0)
        merged = {**vad_params, **seg_cfg} if (backend.startswith("silero") or backend == "b200-vad") else dict(seg_cfg)
        merged.pop("backend", None)
        try:
            self._external_segmenter = self._make_segmenter(backend, merged)
        except Exception as e:  # same failure contract as whisper_pro_asr.py:102-104
            raise ValueError(f"Speech Segmenter not configured - this is an architecture violation: {e}")

        self.whisper_params: Dict = {}
        self.whisper_params.update(decoder_params)
        self.whisper_params.update(provider_params)
        raw_threshold = self.whisper_params.get("logprob_threshold", -1.0)
        self.logprob_threshold = float(raw_threshold) if raw_threshold is not None else None
        self.logprob_margin = float(self.whisper_params.get("logprob_margin", 0.0) or 0.0)
        self.drop_nonverbal_vocals = bool(self.whisper_params.get("drop_nonverbal_vocals", False))
        raw_enabled = self.whisper_params.get("post_model_filter_enabled")
        self.post_model_filter_enabled = True if raw_enabled is None else bool(raw_enabled)
        self._segment_filter = H.LogprobGate(self.post_model_filter_enabled, self.logprob_threshold, self.logprob_margin,
                                             self.drop_nonverbal_vocals)
        for k in ("logprob_margin", "drop_nonverbal_vocals", "post_model_filter_enabled"):
            self.whisper_params.pop(k, None)
        self.whisper_params["task"] = task
        self.task = task
        self._filter_statistics = {"logprob_filtered": 0, "nonverbal_filtered": 0}
        self.suppress_low = ["Thank you", "視聴", "Thanks for"]
        self.suppress_high = ["視聴ありがとうございました", "ご視聴ありがとうございました", "字幕作成者", "提供", "スポンサー"]
        self._last_vad_segments: List[Dict] = []
        self._last_full_results: List[Dict] = []
        from . import model as M
        self.whisper_model = M.load_model(self.model_name, device=self.device, state_dict=model_config.get("state_dict"),
                                          max_batch=int(model_config.get("max_batch", 64)))

    @staticmethod
    def _make_segmenter(backend: str, cfg: Dict):
        if backend == "b200-vad":
            from .segmenter import B200SpeechSegmenter
            return B200SpeechSegmenter(**cfg)
        from whisperjav.modules.speech_segmentation import SpeechSegmenterFactory  # reference factory when present
        return SpeechSegmenterFactory.create(backend, config=cfg)

    # ---- statistics (balanced/fidelity pipelines read these) --------------------------------------
    def reset_statistics(self) -> None:
        self._filter_statistics = {"logprob_filtered": 0, "nonverbal_filtered": 0}

    def get_filter_statistics(self) -> Dict[str, int]:
        return dict(self._filter_statistics)

    def get_last_vad_segments(self) -> List[Dict]:
        return list(self._last_vad_segments)

    # ---- transcription -----------------------------------------------------------------------------
    def _prepare_whisper_params(self) -> Dict:
        p = self.whisper_params.copy()
        if isinstance(p.get("temperature"), list):
            p["temperature"] = tuple(p["temperature"])
        p.setdefault("verbose", None)
        return p

    def _minimal_whisper_params(self) -> Dict:
        return {"task": self.whisper_params.get("task", "transcribe"), "language": self.whisper_params.get("language", "ja"),
                "temperature": 0.0, "beam_size": 3, "fp16": True, "verbose": None}  # whisper_pro_asr.py:446-454

    def _run_batch(self, chunks: List[np.ndarray]) -> List[Optional[dict]]:
        params = self._prepare_whisper_params()
        self.tracer.emit_transcribe_params(params=params, audio_info={"chunks": len(chunks)}, context="b200_transcribe_batch")
        try:
            return self.whisper_model.transcribe_batch(chunks, **params)
        except Exception as e:  # retry once with minimal params, then give up on the batch (whisper_pro_asr.py:432-444)
            logger.error(f"B200 transcription failed: {e}", exc_info=True)
            minimal = self._minimal_whisper_params()
            try:
                logger.warning(f"Retrying transcription with minimal parameters: {minimal}")
                return self.whisper_model.transcribe_batch(chunks, **minimal)
            except Exception as e2:
                logger.error(f"Original error: {e}; fallback error: {e2}")
                return [None] * len(chunks)

    def transcribe(self, audio_path: Union[str, Path], **kwargs) -> Dict:
        self._last_full_results = []
        audio_path = Path(audio_path)
        if "task" in kwargs:
            t = kwargs.pop("task")
            if t != self.task:
                self.whisper_params["task"] = t
                self.task = t
        audio_data, sample_rate = read_wav_mono(audio_path)
        audio_duration = len(audio_data) / sample_rate if sample_rate else 0.0
        result = self._external_segmenter.segment(audio_data, sample_rate=sample_rate)
        vad_segments = result.to_legacy_format()
        self._last_vad_segments = [{"start_sec": round(s["start_sec"], 3), "end_sec": round(s["end_sec"], 3)}
                                   for g in (vad_segments or []) for s in g]
        language = self.whisper_params.get("language", "ja")
        full_clip = False
        if not vad_segments:
            if self._external_segmenter.name == "none" or H.vad_looks_broken(vad_segments, audio_duration):
                full_clip = True
            else:
                return {"segments": [], "text": "", "language": language}
        elif H.vad_looks_broken(vad_segments, audio_duration):
            full_clip = True
        if full_clip:
            spans = [(0.0, len(audio_data) / 16000.0, audio_data)]
        else:
            spans = []
            for g in vad_segments:
                s, e = g[0]["start_sec"], g[-1]["end_sec"]
                spans.append((s, e, audio_data[int(s * sample_rate): int(e * sample_rate)]))
        results = self._run_batch([sp[2] for sp in spans])
        all_segments: List[Dict] = []
        for (s, e, _), res in zip(spans, results):
            if res:
                self._last_full_results.append({"group_start_sec": s, "group_end_sec": e, "result": res})
            if res and res.get("segments"):
                all_segments.extend(self._process_segments(res["segments"], s))
        return {"segments": all_segments, "text": " ".join(x["text"] for x in all_segments), "language": language}

    def _process_segments(self, raw_segments: List[Dict], start_sec: float) -> List[Dict]:
        out: List[Dict] = []
        for seg in raw_segments:
            text = seg.get("text", "").strip()
            if not text:
                continue
            if any(w in text for w in self.suppress_high):
                continue
            avg_logprob = seg.get("avg_logprob", 0.0)
            for w in self.suppress_low:
                if w in text:
                    avg_logprob -= 0.15
            duration = max(0.0, float(seg.get("end", 0.0) - seg.get("start", 0.0)))
            drop, reason, _ = self._segment_filter.should_filter(avg_logprob=avg_logprob, duration=duration, text=text)
            if drop:
                self._filter_statistics["logprob_filtered" if reason == "logprob" else "nonverbal_filtered"] += 1
                continue
            out.append({"start": float(seg.get("start", 0.0)) + start_sec, "end": float(seg.get("end", 0.0)) + start_sec,
                        "text": text, "avg_logprob": avg_logprob})
        return out

    def transcribe_to_srt(self, audio_path: Union[str, Path], output_srt_path: Union[str, Path], **kwargs) -> Path:
        output_srt_path = Path(output_srt_path)
        result = self.transcribe(Path(audio_path), **kwargs)
        output_srt_path.parent.mkdir(parents=True, exist_ok=True)
        with open(output_srt_path, "w", encoding="utf-8") as f:
            f.write(compose_srt(result.get("segments", [])))
        try:
            if self._last_full_results:
                with open(output_srt_path.with_suffix(".transcribe.json"), "w", encoding="utf-8") as f:
                    json.dump(self._last_full_results, f, ensure_ascii=False, indent=2, default=str)
        except Exception as e:  # diagnostic artefact only
            logger.warning(f"Failed to save transcription results JSON (non-fatal): {e}")
        return output_srt_path

    def cleanup(self):
        try:
            if getattr(self, "whisper_model", None) is not None:
                self.whisper_model.close()
                self.whisper_model = None
        finally:
            gc.collect()
