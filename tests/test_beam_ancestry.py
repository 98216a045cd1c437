"""The invariant behind the device beam search's KV handling: resolving a row's history through its
ancestry table gives exactly the cache upstream obtains by permuting the whole self-attention cache every step
(decoding.py::PyTorchInference.rearrange_kv_cache).  Pure bookkeeping, simulated with numpy -- CPU only.

Device rule (csrc/decode.cu::beam_select_kernel, csrc/attention.cu::attn_dec_self_kernel): at step p row r writes its K/V at
physical [r][p]; when row i continues parent s, ``anc_next[i][t] = anc_cur[s][t]`` for t < p and ``anc_next[i][p] = s``; the
reader at step p + 1 takes position t <= p from physical row ``anc[i][t]`` and position p + 1 from itself.  Tables are
double buffered by step parity, as on the device."""
import numpy as np
import pytest


@pytest.mark.parametrize("beam,n_audio,n_initial,steps,seed", [(2, 3, 3, 12, 0), (3, 2, 4, 20, 1), (5, 1, 1, 15, 2), (1, 4, 2, 6, 3)])
def test_ancestry_tables_equal_cache_permutation(beam, n_audio, n_initial, steps, seed):
    rng = np.random.default_rng(seed)
    rows, n_ctx = beam * n_audio, n_initial + steps + 1
    # what a row "computes" at a position depends on its whole token history: model it as a hash of the history
    def kv_of(history):
        return hash(tuple(history)) % (1 << 31)

    hist = [[100 + t for t in range(n_initial)] for _ in range(rows)]       # identical prompts within (and across) windows
    # upstream: a cache tensor [rows][pos] that is permuted along rows every step
    up_cache = np.zeros((rows, n_ctx), dtype=np.int64)
    # device: physical cache written in place + double-buffered ancestry tables pre-filled with the row id
    phys = np.zeros((rows, n_ctx), dtype=np.int64)
    anc = np.repeat(np.arange(rows)[None, :, None], 2, axis=0).repeat(n_ctx, axis=2).astype(np.int16)
    for p in range(n_initial + steps - 1):
        # the decoder step at position p: every row appends the K/V of its current history prefix [0..p]
        for r in range(rows):
            v = kv_of(hist[r][: p + 1])
            up_cache[r, p] = v
            phys[r, p] = v
        # check: what the device reader sees for every row equals upstream's cache row
        cur = anc[p & 1]
        for r in range(rows):
            seen = [phys[cur[r, t], t] for t in range(p)] + [phys[r, p]]
            assert seen == up_cache[r, : p + 1].tolist(), (p, r)
        if p + 1 < n_initial:
            continue                                                          # prompt positions: no selection yet
        # beam selection: within every window each new row picks a parent row of the same window and a new token
        src = np.concatenate([a * beam + rng.integers(0, beam, beam) for a in range(n_audio)])
        new_tok = rng.integers(0, 1000, rows)
        up_cache = up_cache[src].copy()                                       # rearrange_kv_cache(source_indices)
        hist = [hist[src[i]][: p + 1] + [int(new_tok[i])] for i in range(rows)]
        nxt = anc[(p + 1) & 1]
        for i in range(rows):
            nxt[i, :p] = cur[src[i], :p]
            nxt[i, p] = src[i]
