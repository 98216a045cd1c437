"""Minimal WAV / SRT I/O for the boundary classes.  The reference hands stages 16 kHz mono PCM16 WAV
files (whisperjav/modules/audio_extraction.py:48-57, scene_detection_backends/utils.py:106-140) and reads
them with ``soundfile`` (whisper_pro_asr.py:250-252); ``soundfile`` / ``srt`` are used when importable,
otherwise the stdlib ``wave`` module and a local SRT composer with the same output format."""
from __future__ import annotations

import datetime
import wave
from pathlib import Path
from typing import List, Tuple

import numpy as np


def read_wav_mono(path) -> Tuple[np.ndarray, int]:
    """-> (float32 mono in [-1, 1], sample_rate); multi-channel input is averaged (whisper_pro_asr.py:251-252)."""
    try:
        import soundfile as sf  # type: ignore
        data, sr = sf.read(str(path), dtype="float32")
    except ImportError:
        with wave.open(str(path), "rb") as w:
            sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(n)
        if width == 2:
            data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif width == 4:
            data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif width == 1:
            data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        else:
            raise ValueError(f"unsupported WAV sample width {width}")
        if nch > 1:
            data = data.reshape(-1, nch)
    if data.ndim > 1:
        data = np.mean(data, axis=1)
    return np.ascontiguousarray(data, dtype=np.float32), int(sr)


def write_wav_pcm16(path, audio: np.ndarray, sr: int = 16000) -> None:
    q = np.clip(np.round(np.asarray(audio, dtype=np.float64) * 32767.0), -32768, 32767).astype("<i2")
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(q.tobytes())


def _ts(seconds: float) -> str:
    td = datetime.timedelta(seconds=max(0.0, float(seconds)))
    total_ms = int(round(td.total_seconds() * 1000.0))
    h, rem = divmod(total_ms, 3600000)
    m, rem = divmod(rem, 60000)
    s, ms = divmod(rem, 1000)
    return f"{h:02d}:{m:02d}:{s:02d},{ms:03d}"


def compose_srt(segments: List[dict]) -> str:
    """``srt.compose([srt.Subtitle(index, start, end, content)])`` (whisper_pro_asr.py:515-530)."""
    try:
        import srt  # type: ignore
        subs = [srt.Subtitle(index=i, start=datetime.timedelta(seconds=s["start"]), end=datetime.timedelta(seconds=s["end"]),
                             content=s["text"]) for i, s in enumerate(segments, 1)]
        return srt.compose(subs)
    except ImportError:
        out = []
        for i, s in enumerate(segments, 1):
            out.append(f"{i}\n{_ts(s['start'])} --> {_ts(s['end'])}\n{s['text']}\n\n")
        return "".join(out)
