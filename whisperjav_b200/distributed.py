"""Multi-GPU plumbing for the hot path: windows shard across ranks with no data-path
collective (weights replicated, ~3.1 GB fp16 for large-v3); the only exchange is one all-gather of
packed segment records at the end of a job (NCCL over NVLink on the GPU box, gloo in CPU tests).

The reference has no multi-GPU code at all (one CLI process per GPU in a notebook,
notebook/WhisperJAV_kaggle_parallel_edition.ipynb:428-431); stitching by time offset is what makes
the units independent (whisperjav/modules/srt_stitching.py:36-72).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

MAX_TOKENS = 224
STRIDE = 6 + MAX_TOKENS  # int32 words per record


def shard_units(n_units: int, rank: int, world: int, weights: Sequence[float] = None) -> List[int]:
    """Unit ids owned by ``rank``.  With ``weights`` (e.g. VAD speech seconds, a proxy for decode length)
    units are dealt longest-first to the least-loaded rank; otherwise round-robin ``i % world``."""
    if weights is None:
        return [i for i in range(n_units) if i % world == rank]
    order = sorted(range(n_units), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world
    owner = [0] * n_units
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += float(weights[i])
    return [i for i in range(n_units) if owner[i] == rank]


def pack_records(records: Sequence[Tuple[int, float, float, float, float, Sequence[int]]]) -> torch.Tensor:
    """(unit_id, start_s, end_s, avg_logprob, no_speech_prob, tokens) -> int32 [n, STRIDE]."""
    out = np.zeros((len(records), STRIDE), dtype=np.int32)
    for k, (uid, start, end, alp, nsp, toks) in enumerate(records):
        toks = list(toks)[:MAX_TOKENS]
        out[k, 0] = int(uid)
        out[k, 1] = int(round(start * 1000.0))
        out[k, 2] = int(round(end * 1000.0))
        out[k, 3:5] = np.array([alp, nsp], dtype=np.float32).view(np.int32)
        out[k, 5] = len(toks)
        out[k, 6:6 + len(toks)] = toks
    return torch.from_numpy(out)


def unpack_records(t: torch.Tensor) -> List[dict]:
    a = t.cpu().numpy()
    res = []
    for row in a:
        alp, nsp = row[3:5].view(np.float32)
        n = int(row[5])
        res.append({"unit": int(row[0]), "start": row[1] / 1000.0, "end": row[2] / 1000.0, "avg_logprob": float(alp),
                    "no_speech_prob": float(nsp), "tokens": row[6:6 + n].tolist()})
    return res


def gather_segment_records(records: torch.Tensor, device="cuda") -> List[dict]:
    """All-gather every rank's packed records (ragged: counts first, then one padded all-gather) and return
    them sorted by (unit, start) on every rank.  World size 1 is a no-op."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sorted(unpack_records(records), key=lambda r: (r["unit"], r["start"]))
    world = dist.get_world_size()
    dev = torch.device(device)
    n = torch.tensor([records.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    buf = torch.zeros(mx, STRIDE, dtype=torch.int32, device=dev)
    buf[: records.shape[0]] = records.to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)
    out: List[dict] = []
    for r, c in enumerate(counts):
        out.extend(unpack_records(gathered[r][:c]))
    return sorted(out, key=lambda r: (r["unit"], r["start"]))
