"""Seeded audio cases shared by tests/golden/make_scene_kats.py (which runs the reference's driver on them) and the scene tests."""
import numpy as np

CASES = [
    {"name": "film_8min_defaults", "film": (480.0, 11)},
    {"name": "film_5min_legacy_kwargs", "film": (300.0, 12),
     "kwargs": {"max_duration_s": 20.0, "min_duration_s": 0.5, "max_silence": 1.2, "energy_threshold": 30, "pass2_max_silence_s": 0.6,
                "pass2_energy_threshold": 36, "pad_edges_s": 0.25}},
    {"name": "murmur_brute_force", "hand": "murmur"},
    {"name": "murmur_no_fallback", "hand": "murmur", "kwargs": {"brute_force_fallback": False}},
    {"name": "short_tail_block", "hand": "tail"},
    {"name": "all_silence", "hand": "silence"},
    {"name": "one_long_tone", "hand": "tone"},
    {"name": "sr_22050", "hand": "sr22050"},
]


def _tone(rng, seconds, dbfs, sr=16000):
    return (rng.standard_normal(int(seconds * sr)) * 10 ** (dbfs / 20)).astype(np.float32)


def build_case(case):
    """-> (float32 mono audio, sample rate)"""
    if "film" in case:
        from whisperjav_b200.synth import film_audio
        seconds, seed = case["film"]
        return film_audio(seconds, seed), 16000
    kind = case["hand"]
    if kind.startswith("silero_"):
        return _silero_case(kind)
    rng = np.random.default_rng(99)
    if kind == "murmur":   # passes the 32 dB gate, never the 38 dB one, longer than max_duration -> brute-force split
        a = np.concatenate([_tone(rng, 3.0, -80), _tone(rng, 5.0, -20), _tone(rng, 2.5, -80), _tone(rng, 71.3, -55.5), _tone(rng, 2.2, -80),
                            _tone(rng, 0.25, -20), _tone(rng, 2.0, -80)])
        return a, 16000
    if kind == "tail":     # the stream ends inside a block while a token is open (short last block)
        a = np.concatenate([_tone(rng, 1.0, -80), _tone(rng, 12.3456, -25), _tone(rng, 0.4, -80), _tone(rng, 3.21, -25)])
        return a[: len(a) - 137], 16000
    if kind == "silence":
        return _tone(rng, 20.0, -85), 16000
    if kind == "tone":     # one chapter far above both gates and longer than max_duration: pass 2 cuts at max_dur
        return _tone(rng, 95.0, -20), 16000
    if kind == "sr22050":  # block = int(0.05 * 22050) = 1102 samples
        return np.concatenate([_tone(rng, 1.0, -80, 22050), _tone(rng, 8.0, -22, 22050), _tone(rng, 2.4, -80, 22050), _tone(rng, 40.0, -22, 22050),
                               _tone(rng, 1.0, -80, 22050)]), 22050
    raise KeyError(kind)


# ---- Silero-style detector (scene pass 2 by VAD): a FAKE gate shared by the KAT generator and the tests ----------------------------

def fake_vad_probs(x):
    """Per-512-sample-window "speech probability" = clipped window RMS * 4 -- stands in for the VAD network on both sides."""
    x = np.asarray(x, dtype=np.float32)
    n = (len(x) + 511) // 512
    pad = np.zeros(n * 512, dtype=np.float32)
    pad[: len(x)] = x
    rms = np.sqrt(np.mean(pad.reshape(n, 512).astype(np.float64) ** 2, axis=1))
    return np.clip(rms * 4.0, 0.0, 1.0).astype(np.float32)


SILERO_CASES = [
    {"name": "long_chapter_defaults", "hand": "silero_long"},
    {"name": "short_max_duration", "hand": "silero_long", "kwargs": {"max_duration_s": 60.0, "silero_min_silence_ms": 700, "silero_threshold": 0.2,
                                                                     "silero_neg_threshold": 0.1}},
    {"name": "vad_finds_nothing", "hand": "silero_murmur", "kwargs": {"max_duration_s": 40.0}},
]


def _silero_case(kind):
    rng = np.random.default_rng(123)
    parts = []
    if kind == "silero_long":
        # a 9-minute chapter without any pause of 1.8 s at the pass-1 gate (room tone at -50 dBFS keeps it open) whose speech bursts
        # the gate separates by 2-4 s; then a long silence, a 20 s chapter, a long silence, a 70 s chapter
        for _ in range(22):
            parts += [_tone(rng, float(rng.uniform(8.0, 30.0)), -12), _tone(rng, float(rng.uniform(2.0, 4.0)), -50)]
        parts += [_tone(rng, 3.0, -85), _tone(rng, 20.0, -12), _tone(rng, 2.5, -85)]
        for _ in range(4):
            parts += [_tone(rng, float(rng.uniform(10.0, 20.0)), -12), _tone(rng, 2.2, -50)]
        parts += [_tone(rng, 1.0, -85)]
    else:  # above the pass-1 gate throughout, far below the VAD's threshold: pass 2 returns nothing -> brute force at 29 s
        parts += [_tone(rng, 2.0, -85), _tone(rng, 100.0, -52), _tone(rng, 2.5, -85), _tone(rng, 6.0, -12), _tone(rng, 1.0, -85)]
    return np.concatenate(parts), 16000
