"""Host half of the word-timestamp path (whisperjav_b200/timing.py: word grouping, merge_punctuations, long-word heuristics,
segment adjustment, get_end) against the oracle's restatement of openai-whisper timing.py (oracle/timing_oracle.py) on scripted
alignments -- no device, no model: the oracle's DTW / matrix stages are replaced by the same scripted path on both sides."""
import copy

import numpy as np
import pytest

from oracle import timing_oracle as to
from oracle import whisper_oracle as wo
from whisperjav_b200 import model as M
from whisperjav_b200 import timing as TM

EOT, TSB = 50257, 50364


def detok_char(t):
    if t % 7 == 0:
        return "。"
    if t % 11 == 0:
        return "、"
    if t % 13 == 0:
        return "?"
    return chr(0x4E00 + t % 2000)


def decode(ids):
    return "".join("<|endoftext|>" if t == EOT else detok_char(t) for t in ids)


def scripted_case(rng):
    n_seg = int(rng.integers(1, 4))
    seek = int(rng.integers(0, 5)) * 100
    segments, t0 = [], 0
    for _ in range(n_seg):
        n = int(rng.integers(1, 9))
        dur = int(rng.integers(10, 200))
        toks = [TSB + t0] + [int(x) for x in rng.integers(1, 40000, n)] + [TSB + t0 + dur]
        segments.append({"seek": seek, "start": seek / 100 + t0 * 0.02, "end": seek / 100 + (t0 + dur) * 0.02, "tokens": toks,
                         "text": decode([t for t in toks if t < EOT])})
        t0 += dur
    text = [t for s in segments for t in s["tokens"] if t < EOT]
    n_rows = len(text) + 1
    # a monotone DTW path over n_rows rows and 400 frames: row i occupies frames [cut[i], cut[i+1])
    cuts = np.sort(rng.integers(0, 400, n_rows - 1)) if n_rows > 1 else np.array([], int)
    cuts = np.concatenate([[0], cuts, [400]])
    ti, fi = [], []
    for i in range(n_rows):
        span = range(cuts[i], max(cuts[i + 1], cuts[i] + 1))
        for f in span:
            ti.append(i)
            fi.append(min(f, 399))
    probs = rng.random(len(text))
    return segments, text, np.array(ti), np.array(fi), probs


@pytest.mark.parametrize("seed", range(12))
def test_add_word_timestamps_matches_oracle(monkeypatch, seed):
    rng = np.random.default_rng(seed)
    segments, text, ti, fi, probs = scripted_case(rng)
    last_speech = float(rng.uniform(0, 3))
    # ---- oracle side: its own find_alignment / add_word_timestamps with the matrix and DTW stages scripted
    monkeypatch.setattr(to, "alignment_matrix", lambda *a, **k: (np.zeros((len(text) + 1, 400)), list(probs)))
    monkeypatch.setattr(to, "dtw", lambda x: (ti, fi))
    monkeypatch.setattr(to, "_decode_with_eot", lambda toks, eot: decode(toks))

    class FakeMat:  # .double().numpy() of the scripted matrix
        def __neg__(self): return self
        def double(self): return self
        def numpy(self): return np.zeros((len(text) + 1, 400))
    monkeypatch.setattr(to, "alignment_matrix", lambda *a, **k: (FakeMat(), list(probs)))
    ref_segments = copy.deepcopy(segments)
    dims = wo.ModelDimensions(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4)
    to.add_word_timestamps(None, dims, ref_segments, None, 800, language="ja", last_speech_timestamp=last_speech)
    # ---- product side
    M.set_detokenizer(lambda ids: "".join(detok_char(t) for t in ids))
    try:
        tok = M.Tokens(51865, "ja")
        jumps = np.pad(np.diff(ti), (1, 0), constant_values=1).astype(bool)
        alignment = TM.words_from_alignment(text, fi[jumps], probs, M.detokenize_with_specials(tok), "ja", EOT)
        got = copy.deepcopy(segments)
        new_last = TM.add_word_timestamps(got, alignment, EOT, "\"'“¿([{-", "\"'.。,，!！?？:：”)]}、", last_speech)
    finally:
        M.set_detokenizer(None)
    for g, r in zip(got, ref_segments):
        assert g["start"] == r["start"] and g["end"] == r["end"]
        assert g["words"] == r["words"]
    assert TM.last_word_end(got) == to.get_end(ref_segments)
    ends = [s["end"] for s in got if s["words"]]
    assert new_last == (ends[-1] if ends else last_speech)


def test_split_on_spaces_and_merge():
    """English-style grouping: leading-space subwords start words, punctuation is its own word and then merged."""
    table = {1: " Hello", 2: " wor", 3: "ld", 4: ",", 5: " (", 6: "ok", 7: ")", 8: "!"}
    dec = lambda ids: "".join("<|endoftext|>" if t == EOT else table[t] for t in ids)
    words, toks = TM.split_to_word_tokens([1, 2, 3, 4, 5, 6, 7, 8, EOT], dec, "en", EOT)
    assert words == [" Hello", " world", ",", " (ok", ")", "!", "<|endoftext|>"]   # "ok" has no leading space: it joins " ("
    assert toks[1] == [2, 3]
    al = [TM.WordTiming(w, t, i * 0.1, i * 0.1 + 0.1, 1.0) for i, (w, t) in enumerate(zip(words[:-1], toks[:-1]))]
    TM.merge_punctuations(al, "\"'“¿([{-", "\"'.。,，!！?？:：”)]}、")
    assert [a.word for a in al if a.word] == [" Hello", " world,", " (ok)!"]
    assert [a.tokens for a in al if a.word] == [[1], [2, 3, 4], [5, 6, 7, 8]]
