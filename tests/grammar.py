"""Hand-built bitstreams that exercise every token of the block grammar (FormatSpecs.md:35-131, libulc/ulcDecoder.c:82-231),
including the ones no encoder-made golden contains: every decimation layout (also layout 0, "unused" but parseable), one- and
two-nybble window control, bare / extended / silent / lone-Fh sub-block headers, zero runs at 1, 16, 33 and 288, noise runs at 16
and 527 at every level, quantiser changes (also the zero-scale ones, and one right before a stop), both stop codes, decaying tails at the
slowest and the fastest decay, runs that end exactly at the end of their sub-block.  Test infrastructure."""
from __future__ import annotations

import numpy as np

LAYOUTS = [0x0000, 0x0008, 0x0019, 0x0091, 0x012A, 0x01A2, 0x02A1, 0x0A21, 0x123B, 0x12B3, 0x13B2, 0x1B32, 0x23B1, 0x2B31, 0x3B21, 0xB321]
COEF_NYBBLES = [2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14]


def _sub_block(rng, S: int, out: list, loud_q: int):
    """One sub-block of S coefficients appended to the nybble list."""
    h = rng.integers(0, 20)
    if h == 0:
        out += [14, 15]                                   # Eh,Fh: silent
        return
    if h == 1:
        out += [15]                                       # lone Fh: the reference's shift by -2 (scale 0 on x86)
    elif h == 2:
        out += [14, int(rng.integers(0, 15))]             # extended quantiser 14..28 (27, 28: scale 0)
    else:
        out += [int(rng.integers(loud_q, 14))]
    rem = S
    edge = [("z", 1), ("z", 16), ("l", 33), ("l", 288), ("n", 16), ("n", 527)]
    while rem > 0:
        t = rng.integers(0, 100)
        if t < 55:
            out.append(int(rng.choice(COEF_NYBBLES))); rem -= 1
        elif t < 70:
            n = int(min(rem, rng.integers(1, 17))); out += [0, n - 1]; rem -= n
        elif t < 76 and rem >= 33:
            n = int(min(rem, rng.choice([33, 288, int(rng.integers(33, 289))]))); n = max(n, 33)
            out += [1, (n - 33) >> 4, (n - 33) & 15]; rem -= n
        elif t < 86 and rem >= 16:
            n = int(min(rem, rng.choice([16, 527, int(rng.integers(16, 528))])))
            lv = int(rng.integers(0, 8)); m = n - 16
            out += [8, m >> 5, (m >> 1) & 15, (lv << 1) | (m & 1)]; rem -= n
        elif t < 91:
            out += [15, int(rng.integers(0, 14))]         # quantiser change
        elif t < 93:
            out += [15, 14, int(rng.integers(0, 15))]     # extended quantiser change (13, 14 -> scale 0)
        elif t < 96:
            if rng.integers(0, 2): out += [15, int(rng.integers(0, 14))]     # ... right before a stop
            out += [15, 14, 15]; rem = 0                  # zeros to the end
        elif t < 99:
            out += [15, 15, int(rng.integers(0, 16)), *([0, 1], [15, 15], [int(rng.integers(0, 16)), int(rng.integers(0, 16))])[int(rng.integers(0, 3))]]
            rem = 0                                       # decaying noise to the end
        else:
            kind, n = edge[int(rng.integers(0, len(edge)))]
            if n > rem:
                continue
            if kind == "z": out += [0, n - 1]
            elif kind == "l": out += [1, (n - 33) >> 4, (n - 33) & 15]
            else: out += [8, (n - 16) >> 5, ((n - 16) >> 1) & 15, (int(rng.integers(0, 8)) << 1) | ((n - 16) & 1)]
            rem -= n


def block(rng, N: int, n_chan: int, layout: int | None = None) -> np.ndarray:
    """One valid block as bytes."""
    d = int(rng.integers(0, 16)) if layout is None else layout
    out: list = []
    if d == 1 and rng.integers(0, 2):
        out.append(int(rng.integers(0, 8)))               # one-nybble window control: overlap scale only, layout 1 implied
    else:
        out += [8 | int(rng.integers(0, 8)), d]
    loud_q = int(rng.integers(2, 10))
    for _ in range(n_chan):
        pat = LAYOUTS[d]
        while True:
            _sub_block(rng, N >> (pat & 7), out, loud_q)
            pat >>= 4
            if not pat:
                break
    if len(out) & 1:
        out.append(0)
    a = np.array(out, np.uint8)
    return (a[0::2] | (a[1::2] << 4)).astype(np.uint8)


def streams(n_streams: int, n_blocks: int, N: int = 2048, n_chan: int = 2, seed: int = 1):
    """n_streams streams of n_blocks blocks; every layout appears in every stream (block b of stream s uses layout (s + b) % 16 on the
    first pass).  Returns (list of uint8 arrays, list of per-stream block-size arrays)."""
    rng = np.random.default_rng(seed)
    out, sizes = [], []
    for s in range(n_streams):
        blks = [block(rng, N, n_chan, (s + b) % 16 if b < 16 else None) for b in range(n_blocks)]
        out.append(np.concatenate(blks))
        sizes.append(np.array([len(b) for b in blks], np.uint32))
    return out, sizes
