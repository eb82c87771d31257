"""CPU: the N>1 host logic (stream sharding + stats gather) on a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ulc_codec_b200 import shard


def test_stream_range_partitions():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.stream_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]                                   # contiguous, disjoint
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.stream_range(2, 2, 10)


def _worker(rank, world, port, n_streams, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard.stream_range(rank, world, n_streams)
    # a rank's "work": a checksum over its own stream ids, standing in for the per-stream results
    ids = np.arange(a, b, dtype=np.int64)
    st = shard.RankStats(elapsed_ms=10.0 + rank, audio_seconds=60.0 * (b - a), bytes_in=int(ids.sum()), bytes_out=int((ids * 3).sum()),
                         blocks=1294 * (b - a), bad_streams=0)
    allst = shard.gather_stats(st, dist)
    agg = shard.aggregate(allst)
    t = torch.tensor([st.elapsed_ms])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                          # what bench.py does with its timing scalar
    dist.barrier()
    q.put((rank, (a, b), agg, float(t.item())))
    dist.destroy_process_group()


def test_two_rank_gather_over_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 101
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, span0, agg0, t0), (r1, span1, agg1, t1) = res
    assert span0 == (0, 51) and span1 == (51, 101)
    assert agg0 == agg1                                               # every rank sees the same whole-job figures
    assert agg0["elapsed_ms"] == 11.0 and t0 == t1 == 11.0           # time = max over ranks
    assert agg0["audio_seconds"] == 60.0 * n and agg0["blocks"] == 1294 * n
    assert agg0["bytes_in"] == n * (n - 1) // 2                       # the two shards cover every stream exactly once
    assert agg0["ranks"] == 2


def test_single_process_gather_is_identity():
    st = shard.RankStats(elapsed_ms=5.0, audio_seconds=120.0, blocks=10)
    assert shard.gather_stats(st) == [st]
    assert shard.aggregate([st])["xrealtime"] == 120.0 / 5e-3
