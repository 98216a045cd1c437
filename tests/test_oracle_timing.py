"""Word-timestamp arithmetic of the oracle (openai-whisper timing.py restated): median filter, DTW, token alignment."""
import itertools

import numpy as np
import pytest
import torch

from oracle import timing_oracle as to
from oracle import whisper_oracle as wo
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_weights


def test_median_filter_is_a_sliding_median_with_reflect_padding():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 5, 40, generator=g)
    y = to.median_filter(x, 7)
    xp = torch.nn.functional.pad(x, (3, 3, 0, 0), mode="reflect")
    for t in (0, 1, 17, 39):
        assert torch.equal(y[..., t], xp[..., t: t + 7].median(dim=-1).values)
    assert to.median_filter(torch.randn(4, 3), 7).shape == (4, 3)       # shorter than the pad: returned as is
    assert to.median_filter(torch.arange(9.0), 3).tolist() == [1, 1, 2, 3, 4, 5, 6, 7, 7]


def _brute_force_cost(x):
    n, m = x.shape
    best = None

    def rec(i, j, c):
        nonlocal best
        c += x[i, j]
        if i == n - 1 and j == m - 1:
            best = c if best is None else min(best, c)
            return
        if i + 1 < n and j + 1 < m:
            rec(i + 1, j + 1, c)
        if i + 1 < n:
            rec(i + 1, j, c)
        if j + 1 < m:
            rec(i, j + 1, c)

    rec(0, 0, 0.0)
    return best


def test_dtw_finds_the_cheapest_monotone_path():
    rng = np.random.default_rng(3)
    for n, m in [(1, 1), (1, 5), (4, 1), (3, 4), (4, 6), (5, 5)]:
        x = rng.normal(size=(n, m)).astype(np.float32)
        ti, fi = to.dtw(x)
        assert (ti[0], fi[0]) == (0, 0) and (ti[-1], fi[-1]) == (n - 1, m - 1)
        steps = set(zip(np.diff(ti).tolist(), np.diff(fi).tolist()))
        assert steps <= {(1, 1), (1, 0), (0, 1)}
        assert float(x[ti, fi].sum()) == pytest.approx(_brute_force_cost(x), abs=1e-4)


def test_dtw_follows_a_planted_diagonal():
    n, m = 6, 30
    x = np.ones((n, m), dtype=np.float32)
    for i in range(n):
        x[i, 5 * i: 5 * i + 5] = -1.0          # token i is "heard" in frames 5i .. 5i+4
    ti, fi = to.dtw(x)
    first = [int(fi[np.argmax(ti == i)]) for i in range(n)]
    assert first == [0, 5, 10, 15, 20, 25]


def test_token_alignment_is_monotone_and_inside_the_window():
    dims = DIMS["tiny"]
    w = wo.prepare_weights(synth_weights(dims, seed=11), True)
    a = speech_shaped_audio(8.0, 77)
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(a, dims.n_mels, padding=wo.N_SAMPLES), wo.N_FRAMES)[None]
    xa = wo.encoder_forward(w, dims, mel, True)
    o = wo.DecodingOptions(language="ja", without_timestamps=True, sample_len=9)
    text = wo.decode(w, dims, None, o, True, audio_features=xa)[0].tokens
    num_frames = len(a) // wo.HOP_LENGTH
    al = to.find_token_alignment(w, dims, text, xa, num_frames)
    assert [t.token for t in al] == text
    starts, ends = [t.start for t in al], [t.end for t in al]
    assert all(0.0 <= s <= e <= num_frames / 100 + 1e-6 for s, e in zip(starts, ends))
    assert starts == sorted(starts) and ends == sorted(ends)
    assert all(abs(e - s2) < 1e-9 for e, s2 in zip(ends[:-1], starts[1:]))   # a token ends where the next begins
    assert all(0.0 <= t.probability <= 1.0 for t in al)
    assert to.find_token_alignment(w, dims, [], xa, num_frames) == []
