"""CPU restatement of openai-whisper @ c0d2f62 ``whisper/timing.py`` (word-level timestamps: cross-attention alignment by
dynamic time warping).  TEST INFRASTRUCTURE -- groundwork for SURVEY section 8f-2; no device path consumes it yet.

Restated: ``median_filter``, ``dtw`` (``dtw_cpu`` + ``backtrace``), and ``find_alignment`` up to the per-token jump
times.  Upstream then groups tokens into words with the tokenizer (``split_to_word_tokens``); tokenizer assets are
absent offline, so ``find_token_alignment`` returns one timing per *token* (a word = one token), which is the part
that is arithmetic.  Alignment heads: upstream's default when a checkpoint brings no ``alignment_heads`` dump, i.e.
every head of the upper half of the decoder layers (model.py: ``all_heads[n_text_layer // 2:] = True``).

Parity status: unpinned against the absent package; pinned by properties in tests/test_oracle_timing.py."""
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
