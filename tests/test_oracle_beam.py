"""Beam search in the oracle (openai-whisper decoding.py::BeamSearchDecoder / MaximumLikelihoodRanker restated).

The reference package is absent here, so these tests pin the restatement through its own invariants: hand-computed toy
steps, beam_size = 1 == greedy, and agreement of the KV-cache re-ordering with a cache-free recomputation."""
import math

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo
from whisperjav_b200.synth import DIMS, speech_shaped_audio, synth_weights


def test_update_collapses_identical_beams_and_keeps_the_best_live_sequences():
    # vocab 5, eot = 4, two beams of one audio, both still equal to the prompt [0]
    b = wo.BeamSearch(beam_size=2, eot=4)
    tokens = torch.tensor([[0], [0]])
    logits = torch.tensor([[2.0, 1.0, 0.5, -1.0, 0.0]] * 2)
    slp = torch.zeros(2)
    new_tokens, src, completed = b.update(tokens, logits, slp)
    lp = torch.log_softmax(logits[0], -1)
    assert new_tokens.tolist() == [[0, 0], [0, 1]]          # top-3 candidates 0, 1, 2 collapse to one set; the best two live on
    assert src == [1, 1]                                    # the later beam overwrote the source of the shared candidates
    assert slp.tolist() == pytest.approx([lp[0].item(), lp[1].item()])
    assert not completed and b.finished_sequences == [{}]


def test_eot_candidates_finish_and_patience_sets_the_quota():
    b = wo.BeamSearch(beam_size=2, eot=4, patience=1.5)     # max_candidates = round(3.0) = 3
    assert b.max_candidates == 3
    tokens = torch.tensor([[0, 1], [0, 2]])
    slp = torch.tensor([-0.1, -0.2])
    logits = torch.tensor([[0.0, 0.0, -9.0, -9.0, 3.0],      # beam 0: eot first, then tokens 0 / 1 (tie: top-k order)
                           [1.0, -9.0, -9.0, -9.0, 0.5]])    # beam 1: token 0, then eot
    lp = torch.log_softmax(logits, -1)
    new_tokens, src, completed = b.update(tokens, logits, slp)
    fin = b.finished_sequences[0]
    assert (0, 1, 4) in fin and fin[(0, 1, 4)] == pytest.approx(-0.1 + lp[0, 4].item())
    # ranking: (0,1,4) best, then (0,2,0); (0,2,4) ranks before the weak continuations of beam 0 and finishes on the way
    assert new_tokens[0].tolist() == [0, 2, 0] and src[0] == 1
    assert (0, 2, 4) in fin
    assert len(new_tokens) == 2 and not completed            # 2 of 3 finished
    # one more step in which every beam's best candidate is eot fills the quota
    logits2 = torch.tensor([[-9.0, -9.0, -9.0, 0.0, 5.0]] * 2)
    _, _, completed = b.update(new_tokens, logits2, slp)
    assert completed and len(fin) == 3


def test_finalize_tops_up_from_live_beams_and_ranker_normalises_by_length():
    b = wo.BeamSearch(beam_size=2, eot=9)
    b.finished_sequences = [{(1, 2, 9): -1.0}]
    toks, sums = b.finalize(torch.tensor([[[1, 3, 5], [1, 4, 6]]]), torch.tensor([[-3.0, -2.0]]))
    assert [t.tolist() for t in toks[0]] == [[1, 2, 9], [1, 4, 6, 9]] and sums[0] == [-1.0, -2.0]
    # ranker: logprob / length
    cands = [[torch.tensor([5]), torch.tensor([5, 6, 7, 8])]]
    assert wo.rank_maximum_likelihood(cands, [[-1.0, -2.0]]) == [1]             # -1/1 < -2/4
    assert wo.rank_maximum_likelihood(cands, [[-1.0, -2.0]], length_penalty=0.0) == [0]  # penalty 1: raw sums


@pytest.fixture(scope="module")
def tiny():
    dims = DIMS["tiny"]
    w = wo.prepare_weights(synth_weights(dims, seed=11), True)
    wd = wo.ModelDimensions(**dims.__dict__) if not isinstance(dims, wo.ModelDimensions) else dims
    audio = [speech_shaped_audio(6.0, 4000 + i) for i in range(2)]
    mel = torch.stack([wo.pad_or_trim(wo.log_mel_spectrogram(a, wd.n_mels, padding=wo.N_SAMPLES), wo.N_FRAMES) for a in audio])
    xa = wo.encoder_forward(w, wd, mel, True)
    return w, wd, xa


def test_beam_size_one_is_greedy(tiny):
    w, dims, xa = tiny
    for wt in (True, False):
        o = wo.DecodingOptions(language="ja", without_timestamps=wt, max_initial_timestamp=0.0, sample_len=10)
        greedy = wo.decode(w, dims, None, o, True, audio_features=xa)
        beam = wo.decode(w, dims, None, wo.replace(o, beam_size=1), True, audio_features=xa)
        for g, b in zip(greedy, beam):
            # fp16 logits tie exactly now and then; argmax and top-k break such ties differently (upstream's two decoders do too)
            tie = next((i for i, m in enumerate(g.margins) if m == 0.0), None)
            if tie is None:
                assert b.tokens == g.tokens
                assert b.sum_logprob == pytest.approx(g.sum_logprob, abs=1e-4)
            else:
                assert b.tokens[:tie] == g.tokens[:tie]
            assert b.no_speech_prob == pytest.approx(g.no_speech_prob, abs=1e-6)


def test_cache_reordering_matches_cache_free_recomputation(tiny):
    """The beams found with the re-ordered KV cache are the ones a cache-free decoder (full prefix every step) finds."""
    w, dims, xa = tiny
    o = wo.DecodingOptions(language="ja", without_timestamps=True, beam_size=3, patience=1.0, sample_len=7)
    got = wo.decode(w, dims, None, o, True, audio_features=xa)
    tok = wo.SpecialTokens(dims.n_vocab, language="ja", task="transcribe")
    initial = wo.get_initial_tokens(tok, o, dims.n_text_ctx)
    suppress = wo.get_suppress_tokens(tok, o)
    tokens = torch.tensor([initial]).repeat(xa.shape[0], 1).repeat_interleave(3, dim=0)
    rep = xa.repeat_interleave(3, dim=0)
    slp = torch.zeros(tokens.shape[0])
    beam = wo.BeamSearch(3, tok.eot, 1.0)
    for i in range(7):
        logits = wo.decoder_forward(w, dims, tokens, rep, None, True)[:, -1]
        wo.apply_logit_filters(logits, tokens, tok, o, len(initial), suppress, None)
        tokens, _, completed = beam.update(tokens, logits, slp)
        if completed:
            break
    cand, sums = beam.finalize(tokens.reshape(xa.shape[0], 3, -1), slp.reshape(xa.shape[0], 3))
    cand = [[t[len(initial): int((t == tok.eot).nonzero()[0, 0])] for t in s] for s in cand]
    sel = wo.rank_maximum_likelihood(cand, sums)
    for b, r in enumerate(got):
        assert r.tokens == cand[b][sel[b]].tolist()
        assert r.sum_logprob == pytest.approx(sums[b][sel[b]], abs=5e-2)  # fp16 logit quanta (2^-6) over 7 steps
        assert r.avg_logprob == pytest.approx(r.sum_logprob / (len(r.tokens) + 1))


def test_beam_never_scores_below_greedy_prefix_quality(tiny):
    """With patience 1 the returned hypothesis has a length-normalised score at least that of the greedy one whenever the greedy
    sequence finished inside the beam's horizon (it is one of the candidates the beam could keep at every step)."""
    w, dims, xa = tiny
    o = wo.DecodingOptions(language="ja", without_timestamps=True, sample_len=8)
    greedy = wo.decode(w, dims, None, o, True, audio_features=xa)
    beam = wo.decode(w, dims, None, wo.replace(o, beam_size=4), True, audio_features=xa)
    for g, b in zip(greedy, beam):
        assert math.isfinite(b.avg_logprob) and len(b.tokens) <= 8
        if len(g.tokens) < 8 and g.tokens == b.tokens:
            assert b.sum_logprob == pytest.approx(g.sum_logprob, abs=2e-3)


def test_host_finalize_and_ranker_agree_with_the_oracle():
    """whisperjav_b200.hostlogic.beam_finalize_and_rank (the host half of the device beam search) against the oracle's
    BeamSearch.finalize + rank_maximum_likelihood on random finished / live sets."""
    from whisperjav_b200.hostlogic import beam_finalize_and_rank

    rng = np.random.default_rng(5)
    eot, n_initial = 99, 3
    for trial in range(200):
        beam = int(rng.integers(1, 5))
        T = int(rng.integers(4, 9))
        n_fin = int(rng.integers(0, beam + 2))
        prompt = [7, 8, 9]
        finished = []
        seen = set()
        for _ in range(n_fin):
            L = int(rng.integers(1, T - n_initial + 1))
            seq = tuple(prompt + [int(x) for x in rng.integers(10, 90, L)] + [eot])
            if seq in seen:
                continue
            seen.add(seq)
            finished.append((seq, float(-rng.uniform(0.5, 9.0))))
        live_tok = np.array([prompt + [int(x) for x in rng.integers(10, 90, T - n_initial)] for _ in range(beam)])
        live_sum = -rng.uniform(0.5, 9.0, beam)
        lp = None if trial % 3 else float(rng.uniform(0.0, 1.5))
        ids, score = beam_finalize_and_rank(finished, [(live_tok[j], float(live_sum[j])) for j in range(beam)], beam, n_initial, eot, lp)
        b = wo.BeamSearch(beam, eot)
        b.finished_sequences = [dict(finished)]
        toks, sums = b.finalize(torch.tensor(live_tok)[None], torch.tensor(live_sum, dtype=torch.float64)[None])
        cand = [[t[n_initial: int((t == eot).nonzero()[0, 0])] for t in s] for s in toks]
        sel = wo.rank_maximum_likelihood(cand, sums, lp)[0]
        assert ids == cand[0][sel].tolist(), (trial, finished, live_tok, live_sum)
        assert score == pytest.approx(sums[0][sel])


def _toy_logits(prefix, V):
    """A deterministic 'language model' on a toy vocabulary: logits depend on the whole prefix."""
    g = torch.Generator().manual_seed(hash(tuple(prefix)) % (2 ** 31))
    return torch.randn(V, generator=g) * 2.0


@pytest.mark.parametrize("seed", range(6))
def test_beam_search_is_exhaustive_search_when_the_beam_holds_every_prefix(seed):
    """BeamSearch.update / finalize / ranker pinned against brute force: on a toy vocabulary (V = 4 + eot) and horizon 2 a beam of 4
    prunes nothing at step 1 (4 live prefixes), so (a) after every step the live set must be exactly the top-`beam` of ALL
    non-eot prefixes of that length by cumulative log-probability (here: all of them while they fit), (b) every eot-terminated
    sequence of the best `max_candidates` found on the way must carry its exact brute-force score, and (c) the finished list the
    search ends with must be the best `max_candidates` complete hypotheses a brute-force enumeration finds under the same
    "finish when the candidate list is full" rule applied step by step."""
    V, eot, horizon = 5, 4, 2
    beam = 4                       # 4 live prefixes after step 1: every length-2 prefix extends a kept one
    torch.manual_seed(seed)
    prompt = [int(torch.randint(0, 4, (1,)))]
    bs = wo.BeamSearch(beam_size=beam, eot=eot, patience=16.0)    # 64 candidates: the list never fills before the horizon
    tokens = torch.tensor([prompt] * beam)
    slp = torch.zeros(beam)
    # brute force: cumulative log-prob of every sequence
    def score(seq):
        s, p = 0.0, list(prompt)
        for t in seq:
            s += torch.log_softmax(_toy_logits(p, V), -1)[t].item()
            p.append(t)
        return s
    import itertools
    finished_expected = {}
    kept_prev = []
    for step in range(horizon):
        logits = torch.stack([_toy_logits(tokens[r].tolist(), V) for r in range(beam)])
        tokens, src, completed = bs.update(tokens, logits, slp)
        live = {tuple(t.tolist()[len(prompt):]) for t in tokens}
        # all non-eot sequences of this length (identical beams collapse, so the live set is a set of distinct prefixes)
        every = [s for s in itertools.product(range(4), repeat=step + 1)]
        best = sorted(every, key=lambda s: -score(s))[:beam]
        assert live == set(best), (step, live ^ set(best))
        for r in range(len(tokens)):  # cumulative scores carried with the rows
            assert slp[r].item() == pytest.approx(score(tuple(tokens[r].tolist()[len(prompt):])), abs=1e-5)
        # eot-terminated candidates finish only if they rank above the beam-th live candidate of their step
        cands = [p + (t,) for p in (kept_prev if step else [()]) for t in range(V)]
        n_live = 0
        for c in sorted(cands, key=lambda c: -score(c)):
            if c[-1] == eot:
                finished_expected[tuple(prompt) + c] = score(c)
            else:
                n_live += 1
                if n_live == beam:
                    break
        kept_prev = sorted(live)
        # pad the rows back to `beam` (upstream keeps exactly beam rows; duplicates collapse in the next update)
        if len(tokens) < beam:
            reps = (beam + len(tokens) - 1) // len(tokens)
            slp = slp[: len(tokens)].repeat(reps)[:beam].clone()
            tokens = tokens.repeat(reps, 1)[:beam]
    fin = bs.finished_sequences[0]
    assert set(fin) == set(finished_expected)                        # every eot-terminated hypothesis was found ...
    for k, v in fin.items():
        assert v == pytest.approx(finished_expected[k], abs=1e-5)    # ... with its exact score
    toks, sums = bs.finalize(tokens.view(1, beam, -1), slp.view(1, beam))
    pick = wo.rank_maximum_likelihood(toks, sums)[0]
    cands = {tuple(t.tolist()): s for t, s in zip(toks[0], sums[0])}
    best_seq = max(cands, key=lambda k: cands[k] / len(k))
    assert tuple(toks[0][pick].tolist()) == best_seq
