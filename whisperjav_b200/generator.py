"""B3 -- TextGenerator backend "b200-whisper": the decoupled subtitle pipeline's generator protocol
(whisperjav/modules/subtitle_pipeline/protocols.py:68-107) with the semantics of
``AnimeWhisperGenerator`` (generators/anime_whisper.py:216-331): HF-style features (raw audio padded to
30 s, then log-mel), forced prefix ``<sot><ja><transcribe><notimestamps>``, greedy, ``max_new_tokens``.
``generate_batch`` is one device batch (the reference loops serially, anime_whisper.py:323-331)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, List, Optional

import numpy as np

try:
    from whisperjav.modules.subtitle_pipeline.types import TranscriptionResult  # type: ignore
except Exception:
    from dataclasses import dataclass, field

    @dataclass
    class TranscriptionResult:  # subtitle_pipeline/types.py:75-82
        text: str
        language: str
        metadata: dict = field(default_factory=dict)

N_SAMPLES = 480000


class B200WhisperGenerator:
    def __init__(self, model_id: str = "large-v3", device: str = "auto", dtype: str = "auto", no_repeat_ngram_size: int = 0,
                 max_new_tokens: int = 444, state_dict: Optional[dict] = None, max_batch: int = 64, **kwargs: Any):
        if no_repeat_ngram_size:
            raise ValueError("b200-whisper: no_repeat_ngram_size > 0 is not supported (anime path default is 0)")
        self._config = {"model_id": model_id, "device": device, "dtype": dtype, "no_repeat_ngram_size": no_repeat_ngram_size,
                        "max_new_tokens": int(max_new_tokens), "max_batch": int(max_batch)}
        self._state_dict = state_dict
        self._model = None
        self._loaded = False

    @property
    def is_loaded(self) -> bool:
        return self._loaded

    def load(self) -> None:
        if self._loaded:
            return
        from . import model as M
        dev = self._config["device"]
        dev = "cuda" if dev in ("auto", "cuda") else dev
        name = self._config["model_id"]
        name = name if name in M.DIMS else "large-v3"  # HF ids (e.g. litagin/anime-whisper) are large-v3-shaped
        self._model = M.load_model(name, device=dev, state_dict=self._state_dict, max_batch=self._config["max_batch"])
        self._loaded = True

    def unload(self) -> None:
        if self._model is not None:
            self._model.close()
        self._model = None
        self._loaded = False

    def cleanup(self) -> None:
        self.unload()

    def generate(self, audio_path: Path, language: str = "ja", context: Optional[str] = None, **kwargs: Any) -> TranscriptionResult:
        if not self._loaded:
            raise RuntimeError("B200WhisperGenerator.generate() called before load(). Call load() first.")
        return self.generate_batch([audio_path], language=language)[0]

    def generate_batch(self, audio_paths: List[Path], language: str = "ja", contexts: Optional[List[str]] = None,
                       **kwargs: Any) -> List[TranscriptionResult]:
        if not self._loaded:
            raise RuntimeError("B200WhisperGenerator.generate_batch() called before load(). Call load() first.")
        import torch
        from .audioio import read_wav_mono
        from .model import detokenize
        m = self._model
        out: List[TranscriptionResult] = []
        paths = list(audio_paths)
        for c0 in range(0, len(paths), m.max_batch):
            chunk = paths[c0: c0 + m.max_batch]
            host = torch.zeros(len(chunk), N_SAMPLES, dtype=torch.float32)
            for i, p in enumerate(chunk):
                a = p if isinstance(p, np.ndarray) else read_wav_mono(p)[0]
                a = np.asarray(a, dtype=np.float32)[:N_SAMPLES]  # the HF processor truncates / pads to 30 s
                host[i, : len(a)] = torch.from_numpy(a)
            ns = torch.full((len(chunk),), N_SAMPLES, dtype=torch.int32)
            mel = m.log_mel(host.to(m.device), ns.to(m.device), n_frames=3000, layout="time", reflect_total=N_SAMPLES)
            xa = m.encode(mel)
            res = m.decode_features(xa, language="ja", task="transcribe", without_timestamps=True,
                                    sample_len=self._config["max_new_tokens"])
            for p, r in zip(chunk, res):
                out.append(TranscriptionResult(text=r.text, language="ja",
                                               metadata={"generator": "b200-whisper", "audio_path": str(p) if not isinstance(p, np.ndarray) else "<array>",
                                                         "tokens": r.tokens, "avg_logprob": r.avg_logprob}))
        return out
