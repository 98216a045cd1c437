"""CPU restatement of openai-whisper @ c0d2f62 ``whisper/timing.py`` (word-level timestamps: cross-attention alignment by
dynamic time warping).  TEST INFRASTRUCTURE: the checker of csrc/align.cu + whisperjav_b200/timing.py.

Restated: ``median_filter``, ``dtw`` (``dtw_cpu`` + ``backtrace``), ``find_alignment``, ``merge_punctuations``,
``add_word_timestamps`` and ``transcribe.py::get_end``.  Tokenizer assets are absent offline, so words are formed from the
placeholder detokenisation (one code point per token: for Japanese, ``split_tokens_on_unicode`` then makes every token a word);
the closing EOT is its own "word" and is dropped as upstream drops it.  Alignment heads: upstream's default when a checkpoint
brings no ``alignment_heads`` dump, i.e. every head of the upper half of the decoder layers (model.py:
``all_heads[n_text_layer // 2:] = True``).

Parity status: ``median_filter`` and ``dtw`` are pinned to HF transformers' ``_median_filter`` / ``_dynamic_time_warping``
(tests/golden/hf_timing.npz); the rest is unpinned against the absent package and pinned by properties in
tests/test_oracle_timing.py."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import whisper_oracle as wo


def median_filter(x: torch.Tensor, filter_width: int) -> torch.Tensor:
    """timing.py::median_filter: median of width ``filter_width`` along the last dimension, reflect padding."""
    pad_width = filter_width // 2
    if x.shape[-1] <= pad_width:
        return x  # F.pad needs the padding width to be smaller than the input dimension
    ndim = x.ndim
    if ndim <= 2:
        x = x[None, None, :]
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    x = F.pad(x, (filter_width // 2, filter_width // 2, 0, 0), mode="reflect")
    result = x.unfold(-1, filter_width, 1).sort()[0][..., filter_width // 2]
    if ndim <= 2:
        result = result[0, 0]
    return result


def _backtrace(trace: np.ndarray) -> np.ndarray:
    i = trace.shape[0] - 1
    j = trace.shape[1] - 1
    trace[0, :] = 2
    trace[:, 0] = 1
    result = []
    while i > 0 or j > 0:
        result.append((i - 1, j - 1))
        if trace[i, j] == 0:
            i -= 1
            j -= 1
        elif trace[i, j] == 1:
            i -= 1
        elif trace[i, j] == 2:
            j -= 1
        else:
            raise ValueError("Unexpected trace[i, j]")
    return np.array(result)[::-1, :].T


def dtw(x: np.ndarray) -> np.ndarray:
    """timing.py::dtw_cpu: cheapest monotone path through the cost matrix x [N tokens, M frames] from (0, 0) to
    (N-1, M-1) with steps (1,1), (1,0), (0,1); ties prefer (0,1).  Returns [2, path_len] (text indices, time indices)."""
    x = np.asarray(x, dtype=np.float32)
    N, M = x.shape
    cost = np.ones((N + 1, M + 1), dtype=np.float32) * np.inf
    trace = -np.ones((N + 1, M + 1), dtype=np.float32)
    cost[0, 0] = 0
    for j in range(1, M + 1):
        for i in range(1, N + 1):
            c0 = cost[i - 1, j - 1]
            c1 = cost[i - 1, j]
            c2 = cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = x[i - 1, j - 1] + c
            trace[i, j] = t
    return _backtrace(trace)


@dataclass
class TokenTiming:
    token: int
    start: float
    end: float
    probability: float


def alignment_matrix(weights, dims: wo.ModelDimensions, tokens: Sequence[int], xa: torch.Tensor, num_frames: int,
                     sot_len: int, medfilt_width: int = 7, qk_scale: float = 1.0, sim_fp16: bool = True):
    """The token x frame matrix find_alignment feeds to DTW, and the teacher-forced probability of every text token.
    ``tokens`` = [*sot_sequence, no_timestamps, *text_tokens, eot]; xa [1, n_audio_ctx, n_state]."""
    qks: List[torch.Tensor] = []
    t = torch.tensor([list(tokens)])
    logits = wo.decoder_forward(weights, dims, t, xa, None, sim_fp16, cross_qk=qks)[0]
    n_text = len(tokens) - sot_len - 2  # between no_timestamps and eot
    text_tokens = list(tokens[sot_len + 1: sot_len + 1 + n_text])
    eot = tokens[-1]
    sampled_logits = logits[sot_len:, :eot]
    token_probs = sampled_logits.softmax(dim=-1)
    text_token_probs = token_probs[np.arange(n_text), text_tokens].tolist()
    # heads * tokens * frames: every head of the upper half of the layers
    heads = [qks[layer][0, h] for layer in range(dims.n_text_layer // 2, dims.n_text_layer) for h in range(dims.n_text_head)]
    w = torch.stack(heads)
    w = w[:, :, : num_frames // 2]
    w = (w * qk_scale).softmax(dim=-1)
    std, mean = torch.std_mean(w, dim=-2, keepdim=True, unbiased=False)
    w = (w - mean) / std
    w = median_filter(w, medfilt_width)
    matrix = w.mean(dim=0)
    matrix = matrix[sot_len:-1]
    return matrix, text_token_probs


def find_token_alignment(weights, dims: wo.ModelDimensions, text_tokens: Sequence[int], xa: torch.Tensor, num_frames: int,
                         language: str = "ja", task: str = "transcribe", medfilt_width: int = 7, qk_scale: float = 1.0,
                         sim_fp16: bool = True) -> List[TokenTiming]:
    """timing.py::find_alignment with every token its own word: start/end of each text token in seconds (relative to
    the window), from DTW over the negated alignment matrix.  The closing EOT is the last 'word' and is dropped, as
    upstream drops it."""
    if len(text_tokens) == 0:
        return []
    tok = wo.SpecialTokens(dims.n_vocab, language=language, task=task)
    sot = list(tok.sot_sequence)
    tokens = [*sot, tok.no_timestamps, *text_tokens, tok.eot]
    matrix, probs = alignment_matrix(weights, dims, tokens, xa, num_frames, len(sot), medfilt_width, qk_scale, sim_fp16)
    text_indices, time_indices = dtw(-matrix.double().numpy())
    n_words = len(text_tokens) + 1  # + eot
    word_boundaries = np.arange(n_words)  # np.pad(np.cumsum([1] * (n_words - 1)), (1, 0))
    jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
    jump_times = time_indices[jumps] / wo.TOKENS_PER_SECOND
    start_times = jump_times[word_boundaries[:-1]]
    end_times = jump_times[word_boundaries[1:]]
    return [TokenTiming(int(t), float(s), float(e), float(p)) for t, s, e, p in zip(text_tokens, start_times, end_times, probs)]


# ----------------------------------------------------------------------------- words, punctuation, segment adjustment
@dataclass
class WordTiming:
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float


def _decode_with_eot(tokens: Sequence[int], eot: int) -> str:
    return "".join("<|endoftext|>" if t == eot else wo.placeholder_detokenize([t]) for t in tokens)


def split_to_word_tokens(tokens: Sequence[int], eot: int):
    """tokenizer.py::split_to_word_tokens for a language written without spaces (split_tokens_on_unicode) under the placeholder
    detokeniser: every token decodes to valid unicode on its own, so every token is a word."""
    return [_decode_with_eot([t], eot) for t in tokens], [[t] for t in tokens]


def find_alignment(weights, dims: wo.ModelDimensions, text_tokens: Sequence[int], xa: torch.Tensor, num_frames: int, *,
                   language="ja", task="transcribe", medfilt_width: int = 7, qk_scale: float = 1.0, sim_fp16: bool = True) -> List[WordTiming]:
    """timing.py::find_alignment."""
    if len(text_tokens) == 0:
        return []
    tok = wo.SpecialTokens(dims.n_vocab, language=language, task=task)
    sot = list(tok.sot_sequence)
    tokens = [*sot, tok.no_timestamps, *text_tokens, tok.eot]
    matrix, text_token_probs = alignment_matrix(weights, dims, tokens, xa, num_frames, len(sot), medfilt_width, qk_scale, sim_fp16)
    text_indices, time_indices = dtw(-matrix.double().numpy())
    words, word_tokens = split_to_word_tokens(list(text_tokens) + [tok.eot], tok.eot)
    if len(word_tokens) <= 1:
        return []
    word_boundaries = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
    jump_times = time_indices[jumps] / wo.TOKENS_PER_SECOND
    start_times = jump_times[word_boundaries[:-1]]
    end_times = jump_times[word_boundaries[1:]]
    word_probabilities = [np.mean(text_token_probs[i:j]) for i, j in zip(word_boundaries[:-1], word_boundaries[1:])]
    return [WordTiming(word, tokens_, float(start), float(end), float(probability))
            for word, tokens_, start, end, probability in zip(words, word_tokens, start_times, end_times, word_probabilities)]


def merge_punctuations(alignment: List[WordTiming], prepended: str, appended: str):
    """timing.py::merge_punctuations."""
    i = len(alignment) - 2
    j = len(alignment) - 1
    while i >= 0:
        previous = alignment[i]
        following = alignment[j]
        if previous.word.startswith(" ") and previous.word.strip() in prepended:
            following.word = previous.word + following.word
            following.tokens = previous.tokens + following.tokens
            previous.word = ""
            previous.tokens = []
        else:
            j = i
        i -= 1
    i = 0
    j = 1
    while j < len(alignment):
        previous = alignment[i]
        following = alignment[j]
        if not previous.word.endswith(" ") and following.word in appended:
            previous.word = previous.word + following.word
            previous.tokens = previous.tokens + following.tokens
            following.word = ""
            following.tokens = []
        else:
            i = j
        j += 1


def add_word_timestamps(weights, dims: wo.ModelDimensions, segments: List[dict], xa: torch.Tensor, num_frames: int, *, language="ja",
                        task="transcribe", prepend_punctuations: str = "\"\'“¿([{-", append_punctuations: str = "\"\'.。,，!！?？:：”)]}、",
                        last_speech_timestamp: float = 0.0, sim_fp16: bool = True, **kwargs):
    """timing.py::add_word_timestamps."""
    if len(segments) == 0:
        return
    eot = wo.SpecialTokens(dims.n_vocab, language=language, task=task).eot
    text_tokens_per_segment = [[token for token in segment["tokens"] if token < eot] for segment in segments]
    text_tokens = [t for seg in text_tokens_per_segment for t in seg]
    alignment = find_alignment(weights, dims, text_tokens, xa, num_frames, language=language, task=task, sim_fp16=sim_fp16, **kwargs)
    word_durations = np.array([t.end - t.start for t in alignment])
    word_durations = word_durations[word_durations.nonzero()]
    median_duration = np.median(word_durations) if len(word_durations) > 0 else 0.0
    median_duration = min(0.7, float(median_duration))
    max_duration = median_duration * 2
    if len(word_durations) > 0:
        sentence_end_marks = ".。!！?？"
        for i in range(1, len(alignment)):
            if alignment[i].end - alignment[i].start > max_duration:
                if alignment[i].word in sentence_end_marks:
                    alignment[i].end = alignment[i].start + max_duration
                elif alignment[i - 1].word in sentence_end_marks:
                    alignment[i].start = alignment[i].end - max_duration
    merge_punctuations(alignment, prepend_punctuations, append_punctuations)
    time_offset = segments[0]["seek"] * wo.HOP_LENGTH / wo.SAMPLE_RATE
    word_index = 0
    for segment, text_tokens in zip(segments, text_tokens_per_segment):
        saved_tokens = 0
        words = []
        while word_index < len(alignment) and saved_tokens < len(text_tokens):
            timing = alignment[word_index]
            if timing.word:
                words.append(dict(word=timing.word, start=round(time_offset + timing.start, 2), end=round(time_offset + timing.end, 2),
                                  probability=timing.probability))
            saved_tokens += len(timing.tokens)
            word_index += 1
        if len(words) > 0:
            if words[0]["end"] - last_speech_timestamp > median_duration * 4 and (
                words[0]["end"] - words[0]["start"] > max_duration
                or (len(words) > 1 and words[1]["end"] - words[0]["start"] > max_duration * 2)
            ):
                if len(words) > 1 and words[1]["end"] - words[1]["start"] > max_duration:
                    boundary = max(words[1]["end"] / 2, words[1]["end"] - max_duration)
                    words[0]["end"] = words[1]["start"] = boundary
                words[0]["start"] = max(0, words[0]["end"] - max_duration)
            if segment["start"] < words[0]["end"] and segment["start"] - 0.5 > words[0]["start"]:
                words[0]["start"] = max(0, min(words[0]["end"] - median_duration, segment["start"]))
            else:
                segment["start"] = words[0]["start"]
            if segment["end"] > words[-1]["start"] and segment["end"] + 0.5 < words[-1]["end"]:
                words[-1]["end"] = max(words[-1]["start"] + median_duration, segment["end"])
            else:
                segment["end"] = words[-1]["end"]
            last_speech_timestamp = segment["end"]
        segment["words"] = words


def get_end(segments: List[dict]) -> Optional[float]:
    """transcribe.py::get_end."""
    return next((w["end"] for s in reversed(segments) for w in reversed(s["words"])), segments[-1]["end"] if segments else None)
