"""Host half of ``WhisperB200.transcribe``: one iteration of upstream's seek loop (``WhisperB200._advance``: no-speech skip,
timestamp-token slicing, seek advance, prompt reset) against the oracle's restatement on randomised decoder outputs.  CPU only:
``_advance`` is a static method and needs no device."""
import numpy as np
import pytest

from oracle import whisper_oracle as wo
from whisperjav_b200 import model as M


def _random_tokens(rng, tok, kind):
    tsb, n_ts = tok.timestamp_begin, 1500
    text = lambda k: [int(x) for x in rng.integers(100, 40000, k)]  # noqa: E731
    if kind == 0:      # no timestamps at all
        return text(int(rng.integers(1, 12)))
    if kind == 1:      # <t0> text <t1><t1'> text <t2><t2'> ... closed pairs
        out, t = [], int(rng.integers(0, 50))
        for _ in range(int(rng.integers(1, 4))):
            t2 = t + int(rng.integers(1, 200))
            out += [tsb + t] + text(int(rng.integers(1, 6))) + [tsb + min(t2, n_ts)]
            t = min(t2, n_ts)
        return out
    if kind == 2:      # ends with a single timestamp after text (single_timestamp_ending)
        t = int(rng.integers(0, 50))
        t2 = t + int(rng.integers(1, 300))
        return [tsb + t] + text(3) + [tsb + t2, tsb + t2] + text(int(rng.integers(1, 5))) + [tsb + min(t2 + 40, n_ts)]
    if kind == 3:      # one opening timestamp, unfinished segment
        return [tsb + int(rng.integers(0, 100))] + text(int(rng.integers(1, 8)))
    if kind == 4:      # zero-length pair -> emptied segment
        t = int(rng.integers(0, 500))
        return [tsb + t] + text(2) + [tsb + t, tsb + t] + text(2) + [tsb + t + 3, tsb + t + 3]
    return [tsb] + text(4) + [tsb]  # single <|0.00|> ending: duration stays the window


@pytest.mark.parametrize("seed", range(6))
def test_seek_iteration_matches_the_oracle(seed):
    rng = np.random.default_rng(seed)
    for n_vocab in (51865, 51866):
        tok_o = wo.SpecialTokens(n_vocab, language="ja", task="transcribe")
        tok_g = M.Tokens(n_vocab, "ja", "transcribe")
        assert tok_g.timestamp_begin == tok_o.timestamp_begin and tok_g.eot == tok_o.eot
        for trial in range(60):
            tokens = _random_tokens(rng, tok_o, int(rng.integers(0, 6)))
            seek = int(rng.integers(0, 5)) * 700
            segment_size = int(rng.choice([3000, 1234, 600]))
            res = M.DecodingResult(tokens=list(tokens), text="x", avg_logprob=float(rng.uniform(-2.0, -0.1)),
                                   no_speech_prob=float(rng.uniform(0, 1)), temperature=float(rng.choice([0.0, 0.4, 0.8])),
                                   compression_ratio=1.3, language="ja", sum_logprob=-3.0)
            ns_thr, lp_thr = (0.6, -1.0) if trial % 4 else (None, None)
            cond_prev = bool(trial % 2)
            st = {"seek": seek, "all_tokens": [1, 2, 3], "reset": 0, "segments": [{"id": 0}]}
            M.WhisperB200._advance(st, res, tok_g, segment_size, ns_thr, lp_thr, cond_prev)
            # --- oracle: the body of transcribe()'s while loop
            skipped = False
            if ns_thr is not None:
                should_skip = res.no_speech_prob > ns_thr
                if lp_thr is not None and res.avg_logprob > lp_thr:
                    should_skip = False
                skipped = should_skip
            if skipped:
                assert st["seek"] == seek + segment_size and len(st["segments"]) == 1 and st["all_tokens"] == [1, 2, 3]
                continue
            fields = {"temperature": res.temperature, "avg_logprob": res.avg_logprob, "compression_ratio": res.compression_ratio,
                      "no_speech_prob": res.no_speech_prob}
            segs, advance = wo.slice_segments(list(tokens), tok_o, seek, segment_size, fields, detok=M.detokenize)
            assert st["seek"] == seek + advance, (tokens, seek, segment_size)
            got = st["segments"][1:]
            assert len(got) == len(segs)
            for k, (g, o) in enumerate(zip(got, segs)):
                assert g["id"] == 1 + k
                for key in ("seek", "tokens", "text", "temperature", "avg_logprob", "compression_ratio", "no_speech_prob"):
                    assert g[key] == o[key], (key, tokens)
                assert g["start"] == pytest.approx(o["start"], abs=1e-9) and g["end"] == pytest.approx(o["end"], abs=1e-9)
            assert st["all_tokens"] == [1, 2, 3] + [t for s in segs for t in s["tokens"]]
            expect_reset = len(st["all_tokens"]) if (not cond_prev or res.temperature > 0.5) else 0
            assert st["reset"] == expect_reset
