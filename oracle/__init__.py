"""CPU oracle for the WhisperJAV ASR hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in ``whisperjav_b200/`` (the product) may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU baseline.

Parity status: **unpinned at the arithmetic boundary**.  The reference
(meizhong986/WhisperJAV @ f7862f7) delegates all hot-path arithmetic to
third-party packages that are absent from /root/reference *and* from this
container (openai-whisper 20250625 @ c0d2f62, silero-vad, ten-vad), and the
reference's own tests hold no mel / encoder / token golden vectors
(SURVEY.md section 8c).  The restatement below follows the published
algorithm of openai-whisper @ c0d2f62 and is cross-validated against the one
independent implementation that *is* importable here (HF ``transformers``
Whisper, see tests/test_oracle_vs_hf.py and tests/golden/).  Host-side logic
(grouping, padding, filters) *is* pinned by the reference's own known-answer
tests, restated in tests/test_host_kats.py.
"""
