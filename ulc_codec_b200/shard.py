"""Stream sharding across the GPUs of one box.

Streams are independent - no cross-stream state exists anywhere in the codec - so rank r of G simply owns a
contiguous range of streams and there is NO collective in the data path.  The only communication is one
all-gather of a small per-rank statistics record after the timed region.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass


def stream_range(rank: int, world: int, n_streams: int) -> tuple[int, int]:
    """[first, last) streams owned by `rank`: contiguous, disjoint, covering, sizes differ by at most one."""
    if not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad rank/world/n_streams")
    base, extra = divmod(n_streams, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


@dataclass
class RankStats:
    """What each rank reports after a run (all sums are over the rank's own streams)."""
    elapsed_ms: float = 0.0
    audio_seconds: float = 0.0
    bytes_in: int = 0
    bytes_out: int = 0
    blocks: int = 0
    bad_streams: int = 0


def gather_stats(stats: RankStats, dist=None) -> list[RankStats]:
    """All-gather the per-rank records (NCCL on GPUs, gloo in the CPU tests).  Without an initialised process group
    the caller is the only rank."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return [stats]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, asdict(stats))
    return [RankStats(**o) for o in out]


def aggregate(all_stats: list[RankStats]) -> dict:
    """Whole-job figures: time is the MAX over ranks, work is the SUM (bench.py's contract)."""
    ms = max(s.elapsed_ms for s in all_stats)
    audio = sum(s.audio_seconds for s in all_stats)
    return {"elapsed_ms": ms, "audio_seconds": audio, "xrealtime": audio / (ms / 1e3) if ms > 0 else 0.0,
            "bytes_in": sum(s.bytes_in for s in all_stats), "bytes_out": sum(s.bytes_out for s in all_stats),
            "blocks": sum(s.blocks for s in all_stats), "bad_streams": sum(s.bad_streams for s in all_stats),
            "ranks": len(all_stats)}
