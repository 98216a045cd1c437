"""Long-stream pipeline with in-memory stage hand-off: scenes -> VAD gate -> groups -> batched transcription -> segments.

This is the per-scene loop of the reference's balanced / fidelity pipelines (``for scene in scenes: asr.transcribe(scene_wav)``,
whisperjav/pipelines/fidelity_pipeline.py:348, balanced_pipeline.py:483) and the per-group loop inside the ASR wrapper
(whisperjav/modules/whisper_pro_asr.py:306-314) flattened into device batches:

* the stream is cut into scenes: by the two-pass silence detector with its energy gate on the device
  (``scenes.B200SceneDetector.detect``, one kernel launch per pass over the whole stream -- auditok_backend.py:229-567) when a
  ``scene_detector`` is given, else by a fixed-length cut at ``scene_s``;
* every scene goes through the VAD gate in ONE device pass (``B200SpeechSegmenter.segment_batch``) instead of one CPU model
  call per window per scene; VAD sanity fall-back as whisper_pro_asr.py:262-300 (``hostlogic.vad_looks_broken``);
* every VAD group of every scene becomes one clip of ONE ``transcribe_batch`` call (``max_batch`` windows per device pass, longest
  first so a pass's rows finish together); arrays are handed stage to stage -- the reference writes a WAV per scene and re-reads
  it in the ASR wrapper (scene_detection_backends/utils.py:106-140, whisper_pro_asr.py:250-252);
* the wrapper's segment post-filter and the group / scene time offsets are applied on the host (whisper_pro_asr.py:456-503,
  srt_stitching.py:36-72).

Multi-GPU: groups are the shardable unit.  ``shard`` = (rank, world) deals the groups longest-first to the
least-loaded rank by speech seconds (``distributed.shard_units``); no data-path collective; the caller all-gathers the packed
segment records at the end (``distributed.gather_segment_records``).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import hostlogic as H
from .distributed import shard_units

SR = 16000

# config/components/vad/silero.py:105-114 (balanced) and config/components/asr/openai_whisper.py:225-247 (balanced / fidelity)
BALANCED_VAD = dict(threshold=0.28, min_speech_duration_ms=100, min_silence_duration_ms=300, speech_pad_ms=400,
                    chunk_threshold_s=2.5, max_group_duration_s=6.0)
BALANCED_DECODE = dict(task="transcribe", language="ja", beam_size=2, best_of=2, patience=1.2, suppress_blank=True,
                       without_timestamps=False, max_initial_timestamp=0.0, temperature=(0.0,), compression_ratio_threshold=2.4,
                       logprob_threshold=-1.0, no_speech_threshold=0.71, condition_on_previous_text=False)
# main.py:1221-1224 (anime: TEN-style grouping, greedy HF decode, generators/anime_whisper.py:269-279)
ANIME_VAD = dict(style="ten", threshold=0.5, chunk_threshold_s=0.5, max_group_duration_s=5.0)
ANIME_DECODE = dict(task="transcribe", language="ja", without_timestamps=True, temperature=(0.0,), condition_on_previous_text=False,
                    compression_ratio_threshold=None, logprob_threshold=None, no_speech_threshold=None, sample_len=224)


@dataclass
class Unit:
    """One VAD group = one clip handed to the model."""
    stream: int
    scene: int
    start_sample: int       # in the stream
    end_sample: int
    speech_s: float         # sum of its speech segments (the shard weight: a proxy for decode length)


@dataclass
class StreamResult:
    segments: List[dict]
    units: List[Unit]
    stages_s: Dict[str, float] = field(default_factory=dict)
    stats: Dict[str, float] = field(default_factory=dict)


def cut_scenes(n_samples: int, scene_s: float = 29.0) -> List[Tuple[int, int]]:
    step = int(round(scene_s * SR))
    return [(a, min(a + step, n_samples)) for a in range(0, n_samples, step)]


def detect_scene_cuts(audio: np.ndarray, scene_detector) -> List[Tuple[int, int]]:
    """Sample ranges of the detector's scenes (the slices ``save_scene_wav`` would write, auditok_backend.py:440-441)."""
    scenes, _, _ = scene_detector.detect(audio, SR)
    cuts = [(int(s.start_sec * SR), int(s.end_sec * SR)) for s in scenes]
    return [(a, b) for a, b in cuts if b > a]


def vad_units(segmenter, streams: Sequence[np.ndarray], scene_s: float = 29.0, scene_batch: int = 256,
              stream_ids: Optional[Sequence[int]] = None, scene_detector=None, timing: Optional[Dict[str, float]] = None) -> List[Unit]:
    """Scenes of every stream through the VAD gate (``scene_batch`` scenes per device pass) -> one Unit per VAD group."""
    units: List[Unit] = []
    todo = []
    t0 = time.perf_counter()
    for k, audio in enumerate(streams):
        sid = stream_ids[k] if stream_ids is not None else k
        cuts = cut_scenes(len(audio), scene_s) if scene_detector is None else detect_scene_cuts(audio, scene_detector)
        for j, (a, b) in enumerate(cuts):
            todo.append((sid, j, a, b, audio))
    if timing is not None:
        timing["scenes"] = time.perf_counter() - t0
        timing["n_scenes"] = len(todo)
    for c0 in range(0, len(todo), scene_batch):
        chunk = todo[c0: c0 + scene_batch]
        results = segmenter.segment_batch([aud[a:b] for (_, _, a, b, aud) in chunk], sample_rate=SR)
        for (sid, j, a, b, _), res in zip(chunk, results):
            groups = res.to_legacy_format()
            dur = (b - a) / SR
            if H.vad_looks_broken(groups, dur) or (not groups and segmenter.name == "none"):
                units.append(Unit(sid, j, a, b, dur))  # transcribe the whole scene (whisper_pro_asr.py:262-300)
                continue
            for g in groups:
                s0, e0 = int(g[0]["start_sec"] * SR), int(g[-1]["end_sec"] * SR)
                if e0 > s0:
                    units.append(Unit(sid, j, a + s0, min(a + e0, b), float(sum(x["end_sec"] - x["start_sec"] for x in g))))
    return units


def transcribe_units(model, streams: Dict[int, np.ndarray], units: Sequence[Unit], decode: dict, gate: Optional[H.LogprobGate] = None,
                     clip_chunk: int = 2048) -> List[dict]:
    """Units -> segment dicts in stream time.  Units are decoded longest-first (rows of a device pass then finish together)."""
    order = sorted(range(len(units)), key=lambda i: (-(units[i].end_sample - units[i].start_sample), i))
    out: List[dict] = []
    for c0 in range(0, len(order), clip_chunk):
        idx = order[c0: c0 + clip_chunk]
        clips = [streams[units[i].stream][units[i].start_sample: units[i].end_sample] for i in idx]
        results = model.transcribe_batch(clips, **decode)
        for i, res in zip(idx, results):
            u = units[i]
            off = u.start_sample / SR
            for seg in res["segments"]:
                text = seg["text"].strip()
                if not text:
                    continue
                if gate is not None:
                    drop, _, _ = gate.should_filter(avg_logprob=seg["avg_logprob"], duration=max(0.0, seg["end"] - seg["start"]), text=text)
                    if drop:
                        continue
                out.append({"stream": u.stream, "unit": i, "start": seg["start"] + off, "end": seg["end"] + off, "text": text,
                            "avg_logprob": seg["avg_logprob"], "no_speech_prob": seg["no_speech_prob"], "tokens": seg["tokens"]})
    out.sort(key=lambda s: (s["stream"], s["start"]))
    return out


def transcribe_streams(model, segmenter, streams: Sequence[np.ndarray], decode: Optional[dict] = None, scene_s: float = 29.0,
                       sync=None, scene_detector=None) -> StreamResult:
    """The whole path for a list of streams on one GPU.  ``sync`` (e.g. torch.cuda.synchronize) is called at the stage
    boundaries so the per-stage wall times are attributable."""
    decode = dict(BALANCED_DECODE if decode is None else decode)
    t0 = time.perf_counter()
    timing: Dict[str, float] = {}
    units = vad_units(segmenter, streams, scene_s, scene_detector=scene_detector, timing=timing)
    if sync:
        sync()
    t1 = time.perf_counter()
    return _finish(model, streams, units, decode, 0, 1, t0, t1, sync, timing=timing)


def _finish(model, streams, units, decode, rank, world, t0, t1, sync, all_units: Optional[List[Unit]] = None,
            timing: Optional[Dict[str, float]] = None) -> StreamResult:
    pool = all_units if all_units is not None else units
    keep = shard_units(len(pool), rank, world, weights=[u.speech_s for u in pool]) if world > 1 else range(len(pool))
    my_units = [pool[i] for i in keep]
    thr = decode.get("logprob_threshold")
    gate = H.LogprobGate(enabled=thr is not None, logprob_threshold=thr)
    segs = transcribe_units(model, {k: s for k, s in enumerate(streams)}, my_units, decode, gate)
    if sync:
        sync()
    t2 = time.perf_counter()
    audio_s = sum(len(s) for s in streams) / SR
    timing = timing or {}
    scenes_s = float(timing.get("scenes", 0.0))
    return StreamResult(segs, my_units, {"scenes": scenes_s, "vad": t1 - t0 - scenes_s, "transcribe": t2 - t1, "total": t2 - t0},
                        {"streams": len(streams), "audio_s": audio_s, "units": len(my_units), "units_total": len(pool),
                         "scenes": int(timing.get("n_scenes", 0)),
                         "unit_audio_s": sum((u.end_sample - u.start_sample) for u in my_units) / SR,
                         "speech_s": sum(u.speech_s for u in my_units)})


def units_to_tensor(units: Sequence[Unit]):
    import torch
    return torch.tensor([[u.stream, u.scene, u.start_sample, u.end_sample, int(round(u.speech_s * 1000))] for u in units],
                        dtype=torch.int64).reshape(-1, 5)


def units_from_tensor(t) -> List[Unit]:
    return [Unit(int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]) / 1000.0) for r in t.tolist()]


def transcribe_streams_distributed(model, segmenter, streams: Sequence[np.ndarray], decode: Optional[dict] = None, scene_s: float = 29.0,
                                   sync=None, device="cuda", scene_detector=None) -> StreamResult:
    """Config-4 shape under torch.distributed: rank r gates the streams ``s % world == r``, the (tiny) unit lists are all-gathered,
    the units are dealt by speech seconds, every rank transcribes its share.  The caller gathers the segment records."""
    import torch
    import torch.distributed as dist
    decode = dict(BALANCED_DECODE if decode is None else decode)
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    t0 = time.perf_counter()
    mine = [k for k in range(len(streams)) if k % world == rank]
    timing: Dict[str, float] = {}
    units = vad_units(segmenter, [streams[k] for k in mine], scene_s, stream_ids=mine, scene_detector=scene_detector, timing=timing)
    if sync:
        sync()
    all_units = units
    if world > 1:
        dev = torch.device(device)
        t = units_to_tensor(units).to(dev)
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        mx = max(max(int(c) for c in counts), 1)
        buf = torch.zeros(mx, 5, dtype=torch.int64, device=dev)
        buf[: t.shape[0]] = t
        got = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(got, buf)
        all_units = [u for r in range(world) for u in units_from_tensor(got[r][: int(counts[r])].cpu())]
        all_units.sort(key=lambda u: (u.stream, u.start_sample))
    t1 = time.perf_counter()
    return _finish(model, streams, units, decode, rank, world, t0, t1, sync, all_units=all_units, timing=timing)
