"""N>1 host-side path on CPU: world_size-2 gloo, window sharding + the one all-gather of packed records."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from whisperjav_b200.distributed import gather_segment_records, pack_records, shard_units, unpack_records


def test_shard_units_partitions_everything():
    for world in (1, 2, 3, 8):
        owned = [shard_units(37, r, world) for r in range(world)]
        assert sorted(i for o in owned for i in o) == list(range(37))
    w = [5.0, 1.0, 1.0, 1.0, 4.0, 1.0]
    a, b = shard_units(6, 0, 2, w), shard_units(6, 1, 2, w)
    assert sorted(a + b) == list(range(6))
    assert abs(sum(w[i] for i in a) - sum(w[i] for i in b)) <= 1.0  # longest-first deal balances decode work


def test_pack_roundtrip():
    recs = [(7, 0.5, 2.25, -0.75, 0.125, [50365, 11, 12, 50400]), (9, 30.0, 31.5, -1.5, 0.9, [])]
    got = unpack_records(pack_records(recs))
    assert got[0]["tokens"] == [50365, 11, 12, 50400] and got[0]["unit"] == 7 and abs(got[0]["end"] - 2.25) < 1e-9
    assert abs(got[1]["avg_logprob"] + 1.5) < 1e-7 and got[1]["tokens"] == []


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_units(5, rank, world)
    recs = pack_records([(u, float(u), float(u) + 1.0, -0.1 * u, 0.01 * u, list(range(u + 1))) for u in mine])
    allr = gather_segment_records(recs, device="cpu")
    q.put((rank, [(r["unit"], r["tokens"]) for r in allr]))
    dist.destroy_process_group()


def test_gather_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    expect = [(u, list(range(u + 1))) for u in range(5)]
    for _, got in outs:
        assert got == expect  # every rank ends with every record, ordered by unit
