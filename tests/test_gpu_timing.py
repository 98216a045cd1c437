"""Word-level timestamps on the GPU (csrc/align.cu, WhisperB200.align_windows, whisperjav_b200/timing.py) vs the oracle's
restatement of openai-whisper timing.py (oracle/timing_oracle.py; its median filter and DTW are pinned to HF transformers').

* the token x frame matrix (softmax over content frames -> standardise over tokens -> median 7 -> head mean) within 2e-2 abs of the
  oracle's (values are z-scores of order 1; the inputs are fp16 attention scores on both sides);
* DTW on the device == the oracle's DTW *on the device's own matrix* exactly (same fp32 recurrence and tie rule), and the jump
  frames from the two independent matrices agree for >= 75 % of the tokens within one frame (20 ms; a random model's alignment matrix is nearly flat, so paths are
  sensitive; a trained model's is sharp);
* teacher-forced token probabilities within 15 % relative (= a few fp16 quanta of logit difference);
* ``transcribe(word_timestamps=True)``: the oracle follows the device window by window (tokens, segments, words, seeks)."""
import json

import numpy as np
import pytest
import torch

from oracle import parity as P
from oracle import timing_oracle as to
from oracle import whisper_oracle as wo
from whisperjav_b200 import model as M
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_preset, synth_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny():
    dims = DIMS["tiny"]
    w = synth_weights(dims, **synth_preset("tiny"))
    return dims, w, M.WhisperB200(dims, w, max_batch=8), wo.prepare_weights(w, True)


@pytest.fixture(scope="module")
def clips():
    return [speech_shaped_audio(s, 1000 + i) for i, s in enumerate([30.0, 12.0, 5.0, 21.7])]


@pytest.mark.parametrize("mode", ["prefill", "steps"])
def test_alignment_matrix_dtw_and_probs(tiny, clips, diag_dir, mode):
    """Both forms of the teacher-forced pass -- every position a GEMM row (csrc/prefill.cu, the product default) and one position
    per decode step through the step graph -- against the oracle."""
    dims, w, m, pw = tiny
    xa = m.encode(P.gpu_mel(m, clips))
    res = m.decode_features(xa, language="ja", max_initial_timestamp=0.0)
    tok = M.Tokens(dims.n_vocab, "ja")
    text = [[t for t in r.tokens if t < tok.eot] for r in res]
    text[2] = []                                                  # a window without text is skipped
    frames = [min(3000, len(c) // 160) for c in clips]
    got = m.align_windows(xa, text, frames, language="ja", return_matrix=True, mode=mode)
    assert len(got[2][0]) == 0 and len(got[2][1]) == 0
    report = []
    for b in (0, 1, 3):
        jump, probs, mat = got[b]
        sot = list(wo.SpecialTokens(dims.n_vocab, language="ja").sot_sequence)
        tokens = [*sot, tok.no_timestamps, *text[b], tok.eot]
        ref_mat, ref_probs = to.alignment_matrix(pw, dims, tokens, xa[b: b + 1].float().cpu(), frames[b], len(sot))
        assert mat.shape == ref_mat.shape == (len(text[b]) + 1, frames[b] // 2)
        dmat = float((mat - ref_mat).abs().max())
        # the device DTW against the oracle DTW on the SAME (device) matrix: bit-identical recurrence and tie rule
        ti, fi = to.dtw(-mat.double().numpy())
        jumps = np.pad(np.diff(ti), (1, 0), constant_values=1).astype(bool)
        assert np.array_equal(jump, fi[jumps]), b
        # ... and against the oracle end to end
        ti2, fi2 = to.dtw(-ref_mat.double().numpy())
        ref_jump = fi2[np.pad(np.diff(ti2), (1, 0), constant_values=1).astype(bool)]
        close = float(np.mean(np.abs(jump - ref_jump) <= 1))
        dp = float(np.max(np.abs(probs - np.asarray(ref_probs)) / (np.asarray(ref_probs) + 1e-4)))
        report.append({"b": b, "tokens": len(text[b]), "dmatrix": dmat, "jump_within_1_frame": close, "dprob_rel": dp})
        assert dmat <= 2e-2, report[-1]
        assert close >= 0.75, report[-1]   # DTW on a random model's near-flat matrix: paths move under 1e-2 perturbations
        assert dp <= 0.15, report[-1]   # |d log p| = |d logit| up to a few fp16 quanta of a logit ~ 30
    (diag_dir / f"align_tiny_{mode}.json").write_text(json.dumps(report))


def test_prefill_pass_equals_step_pass(tiny, clips):
    """The two passes compute the same thing with different kernels (tcgen05 GEMM over all positions + prefill attention vs the
    decode step graph): matrices agree to fp16 noise, DTW paths almost everywhere."""
    dims, w, m, pw = tiny
    xa = m.encode(P.gpu_mel(m, clips))
    res = m.decode_features(xa, language="ja", max_initial_timestamp=0.0)
    tok = M.Tokens(dims.n_vocab, "ja")
    text = [[t for t in r.tokens if t < tok.eot] for r in res]
    frames = [min(3000, len(c) // 160) for c in clips]
    a = m.align_windows(xa, text, frames, language="ja", return_matrix=True, mode="prefill")
    b = m.align_windows(xa, text, frames, language="ja", return_matrix=True, mode="steps")
    for (ja, pa, ma), (jb, pb, mb) in zip(a, b):
        dmax = float((ma - mb).abs().max())
        assert ma.shape == mb.shape and dmax <= 1e-2
        # DTW is an arg-min over paths of a near-flat matrix (random model): a 1e-2 perturbation may move a stretch of the path
        # (same caveat as the oracle comparison above), so besides the jump frames the check is near-optimality: the prefill
        # pass's path, costed on the step pass's matrix, is within the bound the matrix difference allows of that matrix's optimum
        xa_, xb_ = -ma.double().numpy(), -mb.double().numpy()
        (ta, fa), (tb, fb) = to.dtw(xa_), to.dtw(xb_)
        gap = float(xb_[ta, fa].sum() - xb_[tb, fb].sum())
        assert -1e-9 <= gap <= 2 * max(len(ta), len(tb)) * dmax + 1e-9, (gap, dmax)
        assert np.mean(np.abs(ja - jb) <= 1) >= 0.75
        assert np.max(np.abs(pa - pb) / (pb + 1e-4)) <= 0.1


def test_transcribe_word_timestamps_matches_oracle(tiny, clips, diag_dir):
    """``transcribe(word_timestamps=True)`` end to end, the oracle following the device window by window
    (oracle/parity.py::transcribe_parity): tokens arg-max-or-counted-tie, then the oracle's own add_word_timestamps (its own encoder
    output, its own DTW) on the device's tokens must give the device's segments (bounds within one 20 ms frame), the same number
    of words, >= 90 % of the word boundaries within one frame, and the next seek within one frame."""
    dims, w, m, pw = tiny
    kw = dict(language="ja", task="transcribe", temperature=0.0, no_speech_threshold=0.6, logprob_threshold=-1.0,
              compression_ratio_threshold=2.4, condition_on_previous_text=False, max_initial_timestamp=0.0, word_timestamps=True)
    rep = P.transcribe_parity(m, w, dims, clips, prepared=pw, **kw)
    (diag_dir / "word_timestamps_tiny.json").write_text(json.dumps(rep, indent=1))
    assert rep["ok"], rep["failures"]
    assert rep["words"] >= 100 and rep["words_within_1_frame"] >= 0.9 * rep["words"], rep
    got = m.transcribe_batch(clips, **kw)
    plain = m.transcribe_batch(clips, **{**kw, "word_timestamps": False})
    for g in got:
        for s in g["segments"]:
            assert "words" in s
            for x in s["words"]:
                assert x["end"] >= x["start"] >= 0.0 and 0.0 <= x["probability"] <= 1.0
    # the hook changes what upstream says it changes: segment bounds / seeks, not the decoded ids of the first window
    assert [s["tokens"] for s in got[2]["segments"][:1]] == [s["tokens"] for s in plain[2]["segments"][:1]]
