"""Host logic of the long-stream pipeline (whisperjav_b200/stream.py) on CPU with scripted stand-ins for the two device stages:
scene cut, VAD fall-back, unit construction, longest-first batching, offsets, post-filter, and the world-size-2 sharded run
(gloo) -- the N > 1 path of BASELINE config 4."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from whisperjav_b200 import hostlogic as H
from whisperjav_b200 import stream as S
from whisperjav_b200.distributed import gather_segment_records, pack_records

SR = 16000


class FakeSegmenter:
    """Speech wherever |x| > 0.5 for >= 0.2 s; Silero-style result objects."""
    name = "fake"

    def segment_batch(self, audios, sample_rate=16000, **kw):
        out = []
        for a in audios:
            on = np.abs(a) > 0.5
            edges = np.flatnonzero(np.diff(np.concatenate([[0], on.astype(np.int8), [0]])))
            segs = [H.SpeechSegment(s / SR, e / SR, int(s), int(e)) for s, e in zip(edges[::2], edges[1::2]) if e - s >= 0.2 * SR]
            out.append(H.SegmentationResult(segs, H.group_by_gap(segs, 6.0, 2.5), "fake", len(a) / SR, {}))
        return out


class FakeModel:
    """One segment per clip: text = clip length in samples, avg_logprob from the clip's second sample."""
    def __init__(self):
        self.batches = []

    def transcribe_batch(self, clips, **kw):
        self.batches.append([len(c) for c in clips])
        res = []
        for c in clips:
            lp = -2.0 if c[1] > 0.95 else -0.3
            res.append({"segments": [{"start": 0.1, "end": len(c) / SR, "text": f"n{len(c)}", "avg_logprob": lp, "no_speech_prob": 0.0,
                                      "tokens": [len(c) % 50000]}], "language": "ja", "text": ""})
        return res


def make_stream(seed, seconds=120.0):
    rng = np.random.default_rng(seed)
    a = np.zeros(int(seconds * SR), np.float32)
    t, k = 0, 0
    while t < len(a) - 5 * SR:
        t += int(rng.uniform(0.3, 3.5) * SR)
        n = int(rng.uniform(0.4, 3.0) * SR)
        a[t: t + n] = 0.99 if k % 4 == 3 else 0.9
        t += n
        k += 1
    return a


def test_cut_scenes_covers_the_stream():
    sc = S.cut_scenes(10 * SR + 7, 3.0)
    assert sc[0] == (0, 3 * SR) and sc[-1][1] == 10 * SR + 7 and all(a1 == b0 for (_, b0), (a1, _) in zip(sc, sc[1:]))


def test_units_offsets_and_filter():
    a = make_stream(1)
    m, seg = FakeModel(), FakeSegmenter()
    r = S.transcribe_streams(m, seg, [a], decode=dict(S.BALANCED_DECODE))
    assert r.stats["units"] == len(r.units) > 5 and set(r.stages_s) == {"scenes", "vad", "transcribe", "total"}
    for u in r.units:
        assert np.all(np.abs(a[u.start_sample: u.end_sample][[1, -2]]) > 0.5)        # starts / ends on speech (int(sec * sr) slicing as whisper_pro_asr.py:381)
        assert (u.end_sample - u.start_sample) / SR <= 6.0 + 3.0 + 1e-6                 # group rule: <= max_group + one segment
        assert u.end_sample <= (u.scene + 1) * 29 * SR                                  # never crosses its scene
    # longest first inside the one transcribe call
    assert m.batches[0] == sorted(m.batches[0], reverse=True)
    # every kept segment sits at its unit's offset; the low-logprob units were filtered (logprob_threshold -1.0)
    kept = {s["unit"] for s in r.segments}
    for s in r.segments:
        u = r.units[s["unit"]]
        assert abs(s["start"] - (u.start_sample / SR + 0.1)) < 1e-9 and s["text"] == f"n{u.end_sample - u.start_sample}"
    dropped = [i for i, u in enumerate(r.units) if a[u.start_sample + 1] > 0.95]
    assert dropped and not (set(dropped) & kept) and len(kept) == len(r.units) - len(dropped)
    assert [s["start"] for s in r.segments] == sorted(s["start"] for s in r.segments)


def test_silent_long_scene_falls_back_to_whole_scene():
    """vad_failover.py:26-57: a >= 120 s clip with no VAD segments is transcribed whole."""
    a = np.zeros(130 * SR, np.float32)
    units = S.vad_units(FakeSegmenter(), [a], scene_s=130.0)
    assert len(units) == 1 and (units[0].start_sample, units[0].end_sample) == (0, 130 * SR)
    assert S.vad_units(FakeSegmenter(), [a], scene_s=29.0) == []                        # short silent scenes are skipped


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    streams = [make_stream(10 + k, 90.0) for k in range(3)]
    r = S.transcribe_streams_distributed(FakeModel(), FakeSegmenter(), streams, decode=dict(S.ANIME_DECODE), device="cpu")
    recs = pack_records([(s["stream"] * 100000 + int(s["start"] * 100), s["start"], s["end"], s["avg_logprob"], s["no_speech_prob"], s["tokens"])
                         for s in r.segments])
    allr = gather_segment_records(recs, device="cpu")
    q.put((rank, r.stats["units"], r.stats["units_total"], r.stats["speech_s"], [(x["unit"], x["tokens"]) for x in allr]))
    dist.destroy_process_group()


def test_sharded_run_world2_gloo_matches_single():
    streams = [make_stream(10 + k, 90.0) for k in range(3)]
    single = S.transcribe_streams(FakeModel(), FakeSegmenter(), streams, decode=dict(S.ANIME_DECODE))
    expect = sorted((s["stream"] * 100000 + int(s["start"] * 100), s["tokens"]) for s in single.segments)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    (_, n0, tot0, sp0, all0), (_, n1, tot1, sp1, all1) = outs
    assert tot0 == tot1 == single.stats["units"] and n0 + n1 == tot0 and min(n0, n1) > 0
    assert abs(sp0 - sp1) <= 0.1 * (sp0 + sp1)                                          # dealt by speech seconds: balanced
    assert all0 == all1 == expect                                                       # every rank ends with every record


class NumpySceneDetector:
    """The product's two-pass driver (whisperjav_b200.scenes.two_pass) with the oracle's numpy twin of the energy kernel."""

    def detect(self, audio, sample_rate=16000):
        from oracle import scene_oracle as SO
        from whisperjav_b200 import scenes as SC
        return SC.two_pass(SC.SceneConfig(), len(audio), sample_rate, lambda regions, window: SO.window_sumsq(audio, regions, window))


def test_stream_takes_its_scenes_from_the_detector():
    """The silence detector's scenes replace the fixed 29 s cut: units never cross a detected scene, nothing of a silent gap is
    transcribed, and the scenes are the ones the detector reports."""
    a = np.zeros(int(200 * SR), np.float32)
    rng = np.random.default_rng(4)
    # three chapters (8 s, 65 s with 1.2 s pauses inside, 20 s) separated by > 1.8 s of digital silence
    spans = [(5.0, 13.0), (17.0, 82.0), (90.0, 110.0)]
    for s0, s1 in spans:
        a[int(s0 * SR): int(s1 * SR)] = 0.9
    for p0 in (30.0, 47.5, 66.0):
        a[int(p0 * SR): int((p0 + 1.2) * SR)] = 0.0
    det = NumpySceneDetector()
    scenes, story, _ = det.detect(a, SR)
    assert [(round(x, 2), round(y, 2)) for x, y in story] == [(5.0, 13.0), (17.0, 82.0), (90.0, 110.0)]
    assert len(scenes) == 6 and all(s.end_sec - s.start_sec <= 29.0 for s in scenes)
    m, seg = FakeModel(), FakeSegmenter()
    r = S.transcribe_streams(m, seg, [a], decode=dict(S.BALANCED_DECODE), scene_detector=det)
    assert r.stats["scenes"] == 6 and r.stages_s["scenes"] > 0
    cuts = [(int(s.start_sec * SR), int(s.end_sec * SR)) for s in scenes]
    for u in r.units:
        assert any(c0 <= u.start_sample and u.end_sample <= c1 for c0, c1 in cuts)
        assert np.all(np.abs(a[u.start_sample: u.end_sample][[1, -2]]) > 0.5)
    assert sum(u.end_sample - u.start_sample for u in r.units) <= sum(c1 - c0 for c0, c1 in cuts)
