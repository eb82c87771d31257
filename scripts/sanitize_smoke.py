#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): a few streams through every kernel family,
with CTA reuse (more streams than the persistent grids would need is not possible under the tool's slowdown, so the
ULCB_WS_PERSM / grid caps are left alone and the stream count is kept small but > 1 per kernel)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import wavecheck
from ulc_codec_b200 import capi, corpus

what = sys.argv[1] if len(sys.argv) > 1 else "all"
streams, nbs, bbs = wavecheck.distinct_streams(6, 3, 4)
if what in ("all", "decode"):
    with capi.Context(44100, 2, 2048, max_blocks=8) as ctx:
        kw = {"block_bytes": bbs} if getattr(capi, "HAS_BLOCK_INDEX", False) else {}
        pcm, st, _ = ctx.decode_streams(streams, nbs, fmt="pcm16", **kw)
        assert (st == -1).all()
        pcm, st, _ = ctx.decode_streams(streams, nbs, fmt="f32")
        assert (st == -1).all()
    # the hand-built grammar vectors (every token and decimation layout, decaying tails) through the block-parallel pipeline
    # (token-automaton walk with TMA rings when no index is given) and a multichannel / small-block shape through the generic kernel
    from tests import grammar
    gst, gsz = grammar.streams(5, 6, 2048, 2, seed=3)
    with capi.Context(44100, 2, 2048, max_blocks=6) as ctx:
        for kw in ({}, {"block_bytes": gsz}):
            pcm, st, _ = ctx.decode_streams(gst, 6, fmt="f32", **kw)
            assert (st == -1).all()
    gst, gsz = grammar.streams(3, 5, 512, 3, seed=9)
    with capi.Context(44100, 3, 512, max_blocks=5) as ctx:
        pcm, st, _ = ctx.decode_streams(gst, 5, fmt="pcm16")
        assert (st == -1).all()
if what in ("all", "encode"):
    pcms = [corpus.generate(900 + i, 3 * 2048 - 100, "transient" if i & 1 else "music") for i in range(3)]
    with capi.Context(44100, 2, 2048) as ctx:
        r = ctx.encode_streams(pcms, "vbr", 50.0)
        assert (r["status"] == 0).all()
        r = ctx.encode_streams(pcms[:2], "cbr", 96.0)
        assert (r["status"] == 0).all()
    with capi.Context(44100, 3, 1024) as ctx:                          # run-time channel count, global workspace
        r = ctx.encode_streams([np.concatenate([p, p[::-1]], axis=1)[:4000, :3].copy() for p in pcms[:2]], "vbr", 60.0)
        assert (r["status"] == 0).all()
    with capi.Context(44100, 2, 512) as ctx:
        r = ctx.encode_streams([p[:2000] for p in pcms[:2]], "vbr", 60.0)
        dec, st, _ = ctx.decode_streams(r["streams"], r["block_bytes"].shape[1], fmt="pcm16")
        assert (st == -1).all()
print("sanitize_smoke ok", what)
