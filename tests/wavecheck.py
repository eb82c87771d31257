"""Launch-shape parity checks: push several WAVES of streams through the persistent kernels and compare EVERY
stream with the oracle (test infrastructure; used by tests/test_waves_gpu.py, bench.py's post-run verification and,
as `python -m tests.wavecheck decode ...`, by subprocesses that select a kernel variant through the environment).

A kernel whose CTAs walk more than one stream must reset every piece of per-stream state (pending tail,
LastSubBlockSize, noise-draw position, double-buffer parity, look-ahead FIFOs) at a stream boundary.  The checks
below make stream i on CTA-reuse round r differ from its predecessor on the same CTA: a few dozen distinct streams
(long blocks, window-switched blocks, odd and even block counts), permuted over thousands of slots.
"""
from __future__ import annotations

import sys

import numpy as np

from oracle import oracle
from tests.fixtures import load_golden_stream

N, NCH = 2048, 2


def block_pool(names=("music_cbr128_bs2048", "transient_vbr50_bs2048", "music_vbr75_bs2048")):
    """Reference-encoded blocks of several golden streams (any sequence of valid blocks is a valid stream)."""
    pool = []
    for nm in names:
        g = load_golden_stream(nm)
        offs = np.concatenate([[0], np.cumsum(g["block_bytes"])]).astype(np.int64)
        pool += [g["data"][offs[i]:offs[i + 1]] for i in range(2, len(offs) - 1)]
    return pool


def distinct_streams(n_distinct: int, min_blocks: int, max_blocks: int, seed: int = 1, pool=None):
    """n_distinct streams of min..max blocks drawn from the mixed pool.  Returns (streams, n_blocks, block_bytes)."""
    pool = pool or block_pool()
    rng = np.random.default_rng(seed)
    streams, nbs, bbs = [], [], []
    for d in range(n_distinct):
        nb = min_blocks + d % (max_blocks - min_blocks + 1)
        idx = rng.integers(0, len(pool), nb)
        streams.append(np.concatenate([pool[i] for i in idx]))
        nbs.append(nb)
        bbs.append(np.array([len(pool[i]) for i in idx], np.uint32))
    return streams, nbs, bbs


def slot_order(n_slots: int, n_distinct: int):
    """distinct-stream index of every slot: neighbours differ, and slot s + k*grid differs from slot s for the grids in use."""
    return [(s * 7 + s // n_distinct) % n_distinct for s in range(n_slots)]


def check_decode_waves(capi, n_slots: int = 2300, n_distinct: int = 41, min_blocks: int = 5, max_blocks: int = 7,
                       fmt: str = "pcm16", with_index: bool = False, pipeline=None, pinned: bool = False):
    """Decode n_slots streams in ONE call and compare every one with the oracle.  Returns the number of streams verified."""
    streams, nbs, bbs = distinct_streams(n_distinct, min_blocks, max_blocks)
    refs = []
    for st, nb in zip(streams, nbs):
        f, bits = oracle.Decoder(NCH, N).decode_stream(st, nb)
        assert (bits != 0).all()
        refs.append((oracle.float_to_pcm16(f) if fmt == "pcm16" else f, bits))
    order = slot_order(n_slots, n_distinct)
    batch = [streams[o] for o in order]
    nblk = [nbs[o] for o in order]
    with capi.Context(44100, NCH, N, max_blocks=max_blocks) as ctx:
        kw = {}
        if pipeline:
            ctx.set_decode_pipeline(*pipeline)                     # (chunks, overlap): chunked upload / walk / parse beside the transform
        if pinned:
            # page-locked PCM sink: the kernel stores into it over PCIe and the library pipelines the call on its own
            import torch
            keep = torch.empty(n_slots * max_blocks * N * NCH, dtype=torch.int16 if fmt == "pcm16" else torch.float32).pin_memory()
            kw["out"] = keep.numpy()
        if with_index:
            kw["block_bytes"] = [bbs[o] for o in order]
        pcm, status, bits = ctx.decode_streams(batch, nblk, fmt=fmt, want_bits=True, **kw)
    assert (status == -1).all(), np.nonzero(status != -1)[0][:10]
    for s, o in enumerate(order):
        ref, rbits = refs[o]
        nb = nbs[o]
        got = pcm[s, :nb * N]
        assert np.array_equal(bits[s, :nb].astype(np.int64), rbits.astype(np.int64)), ("bits", s, o)
        if fmt == "pcm16":
            d = np.abs(got.astype(np.int32) - ref.astype(np.int32)).max()
            assert d <= 1, ("pcm16", s, o, int(d))
        else:
            err = np.abs(got.astype(np.float64) - ref).max() / max(1e-30, np.abs(ref).max())
            assert err < 1e-5, ("f32", s, o, float(err))
        assert not pcm[s, nb * N:].any(), ("tail written", s)
    return n_slots


def check_encode_waves(capi, n_slots: int = 450, n_distinct: int = 30, mode: str = "vbr", rate: float = 50.0):
    """Encode n_slots streams (5 blocks each) in ONE call; bytes, block sizes and window decisions of every stream against
    the oracle encoder."""
    from ulc_codec_b200 import corpus
    ns = 3 * N - 300                                                       # -> ceil(n/N)+2 = 5 blocks
    pcms = [corpus.generate(7000 + d, ns, "transient" if d % 3 == 0 else "music") for d in range(n_distinct)]
    refs = [oracle.Encoder(44100, NCH, N).encode_stream(p, mode, rate, 0.7) for p in pcms]
    order = slot_order(n_slots, n_distinct)
    with capi.Context(44100, NCH, N) as ctx:
        r = ctx.encode_streams([pcms[o] for o in order], mode, rate, 0.7)
    assert (r["status"] == 0).all()
    for s, o in enumerate(order):
        ref = refs[o]
        assert np.array_equal(r["block_bytes"][s], ref["block_bytes"]), ("block_bytes", s, o)
        assert np.array_equal(r["window_ctrl"][s], ref["window_ctrl"]), ("window_ctrl", s, o)
        assert np.array_equal(r["streams"][s], ref["data"]), ("bytes", s, o)
    return n_slots


if __name__ == "__main__":
    from ulc_codec_b200 import capi as _capi
    what = sys.argv[1]
    if what == "decode":
        n = check_decode_waves(_capi, int(sys.argv[2]), fmt=sys.argv[3], with_index=len(sys.argv) > 4 and sys.argv[4] == "index")
    else:
        n = check_encode_waves(_capi, int(sys.argv[2]), mode=sys.argv[3], rate=float(sys.argv[4]))
    print(f"wavecheck {what} ok: {n} streams verified")
