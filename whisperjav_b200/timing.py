"""Word-level timestamps: the host half of openai-whisper ``whisper/timing.py`` (``word_timestamps=True`` in every reference preset,
whisperjav/config/components/asr/openai_whisper.py:213,247,281; it moves the segment start / end that land in the SRT and the
seek of the window loop).

The arithmetic -- teacher-forced decoder pass, cross-attention scores of the alignment heads, softmax / standardise / median
filter / head mean, DTW -- runs on the device (csrc/align.cu, ``WhisperB200.align_windows``).  What is left here is list
bookkeeping with upstream's exact rules: token -> word grouping (``tokenizer.split_to_word_tokens``), ``merge_punctuations``,
the long-word truncation heuristics and the segment adjustments of ``add_word_timestamps``.
"""
from __future__ import annotations

import string
from dataclasses import dataclass
from typing import Callable, List, Sequence, Tuple

import numpy as np

TOKENS_PER_SECOND = 50
HOP_LENGTH, SAMPLE_RATE = 160, 16000
NO_SPACE_LANGUAGES = {"zh", "ja", "th", "lo", "my", "yue"}


@dataclass
class WordTiming:
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float


def split_tokens_on_unicode(tokens: Sequence[int], decode: Callable[[Sequence[int]], str]) -> Tuple[List[str], List[List[int]]]:
    """tokenizer.py::split_tokens_on_unicode: cut wherever the tokens so far decode to valid unicode."""
    decoded_full = decode(tokens)
    replacement_char = "�"
    words, word_tokens, current = [], [], []
    unicode_offset = 0
    for token in tokens:
        current.append(token)
        decoded = decode(current)
        if replacement_char not in decoded or decoded_full[unicode_offset + decoded.index(replacement_char)] == replacement_char:
            words.append(decoded)
            word_tokens.append(current)
            current = []
            unicode_offset += len(decoded)
    return words, word_tokens


def split_tokens_on_spaces(tokens: Sequence[int], decode, eot: int) -> Tuple[List[str], List[List[int]]]:
    """tokenizer.py::split_tokens_on_spaces."""
    subwords, subword_tokens_list = split_tokens_on_unicode(tokens, decode)
    words, word_tokens = [], []
    for subword, subword_tokens in zip(subwords, subword_tokens_list):
        special = subword_tokens[0] >= eot
        with_space = subword.startswith(" ")
        punctuation = subword.strip() in string.punctuation
        if special or with_space or punctuation or len(words) == 0:
            words.append(subword)
            word_tokens.append(subword_tokens)
        else:
            words[-1] = words[-1] + subword
            word_tokens[-1].extend(subword_tokens)
    return words, word_tokens


def split_to_word_tokens(tokens: Sequence[int], decode, language: str, eot: int):
    """tokenizer.py::split_to_word_tokens: languages written without spaces split on unicode boundaries, the rest on spaces."""
    if language in NO_SPACE_LANGUAGES:
        return split_tokens_on_unicode(tokens, decode)
    return split_tokens_on_spaces(tokens, decode, eot)


def words_from_alignment(text_tokens: Sequence[int], jump_frames: Sequence[int], token_probs: Sequence[float], decode, language: str,
                         eot: int) -> List[WordTiming]:
    """The tail of timing.py::find_alignment: ``jump_frames[i]`` = frame at which the DTW path enters row i (rows = text tokens
    + the closing EOT), ``token_probs[i]`` = teacher-forced probability of text token i."""
    if len(text_tokens) == 0:
        return []
    words, word_tokens = split_to_word_tokens(list(text_tokens) + [eot], decode, language, eot)
    if len(word_tokens) <= 1:
        return []
    word_boundaries = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    jump_times = np.asarray(jump_frames, dtype=np.float64) / TOKENS_PER_SECOND
    start_times = jump_times[word_boundaries[:-1]]
    end_times = jump_times[word_boundaries[1:]]
    probs = np.asarray(token_probs, dtype=np.float64)
    word_probabilities = [float(np.mean(probs[i:j])) for i, j in zip(word_boundaries[:-1], word_boundaries[1:])]
    return [WordTiming(w, list(t), float(s), float(e), p)
            for w, t, s, e, p in zip(words, word_tokens, start_times, end_times, word_probabilities)]


def merge_punctuations(alignment: List[WordTiming], prepended: str, appended: str) -> None:
    """timing.py::merge_punctuations (in place)."""
    i = len(alignment) - 2
    j = len(alignment) - 1
    while i >= 0:
        previous, following = alignment[i], alignment[j]
        if previous.word.startswith(" ") and previous.word.strip() in prepended:
            following.word = previous.word + following.word
            following.tokens = previous.tokens + following.tokens
            previous.word = ""
            previous.tokens = []
        else:
            j = i
        i -= 1
    i, j = 0, 1
    while j < len(alignment):
        previous, following = alignment[i], alignment[j]
        if not previous.word.endswith(" ") and following.word in appended:
            previous.word = previous.word + following.word
            previous.tokens = previous.tokens + following.tokens
            following.word = ""
            following.tokens = []
        else:
            i = j
        j += 1


def add_word_timestamps(segments: List[dict], alignment: List[WordTiming], eot: int, prepend_punctuations: str, append_punctuations: str,
                        last_speech_timestamp: float) -> float:
    """timing.py::add_word_timestamps after ``find_alignment``: attaches ``words`` to every segment of one window, applies the
    long-word heuristics and moves segment start / end onto the word boundaries.  Returns the updated last_speech_timestamp."""
    if len(segments) == 0:
        return last_speech_timestamp
    text_tokens_per_segment = [[t for t in s["tokens"] if t < eot] for s in segments]
    word_durations = np.array([t.end - t.start for t in alignment])
    word_durations = word_durations[word_durations.nonzero()]
    median_duration = float(np.median(word_durations)) if len(word_durations) > 0 else 0.0
    median_duration = min(0.7, median_duration)
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
    time_offset = segments[0]["seek"] * HOP_LENGTH / SAMPLE_RATE
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
                    or (len(words) > 1 and words[1]["end"] - words[0]["start"] > max_duration * 2)):
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
    return last_speech_timestamp


def last_word_end(segments: List[dict]):
    """transcribe.py::get_end."""
    for s in reversed(segments):
        for w in reversed(s.get("words", [])):
            return w["end"]
    return segments[-1]["end"] if segments else None
