"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, scalar build).

Run in the build container (needs /root/reference -> `make -C oracle ref`):
    python tests/golden/make_golden.py
Each fixture holds: the synthetic-corpus recipe (kind, stream id, seconds - the PCM itself is
regenerated from ulc_codec_b200/corpus.py), the reference encoder's bytes + per-block metadata,
and for the reference decoder's float output a SHA-256 plus head/tail slices.  The oracle
restatement must reproduce the SHA bit-for-bit (tests/test_oracle_golden.py), which pins it on
boxes where oracle/_ref is absent.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbin  # noqa: E402
from ulc_codec_b200 import corpus  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RATE = 44100

#        name                        kind         id  secs  mode   param  bs
CASES = [
    ("music_cbr128_bs2048",          "music",      0, 20.0, "cbr", 128.0, 2048),
    ("music_vbr50_bs2048",           "music",      1,  6.0, "vbr",  50.0, 2048),
    ("music_vbr75_bs2048",           "music",      2,  6.0, "vbr",  75.0, 2048),
    ("transient_vbr50_bs2048",       "transient",  3,  6.0, "vbr",  50.0, 2048),
    ("transient_abr64_bs256",        "transient",  4,  3.0, "abr",  64.0, 256),
    ("transient_abr64_bs8192",       "transient",  5,  8.0, "abr",  64.0, 8192),
    ("music_cbr128_bs512",           "music",      6,  3.0, "cbr", 128.0, 512),
    ("music_vbr60_bs4096",           "music",      7,  6.0, "vbr",  60.0, 4096),
]


def main():
    for name, kind, sid, secs, mode, param, bs in CASES:
        pcm = corpus.generate(sid, int(secs * RATE), kind)
        avgc = 0.0
        if mode == "abr":   # two-pass ABR as the tool is used (README.md:24): complexity from a first run
            first = refbin.encode(pcm, "cbr", param, block_size=bs)
            avgc = float(np.float32(first.complexity.astype(np.float64).mean()))
        enc = refbin.encode(pcm, mode, param, avg_complexity=avgc, block_size=bs, dump=True)
        dec, bits = refbin.decode(enc.data, len(enc.block_bytes), block_size=bs)
        assert np.array_equal((bits + 7) // 8, enc.block_bytes.astype(np.int64))
        # criterion-2 goldens: hashes of the reference's normalised MDCT coefficients and ranks
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            kind=kind, stream_id=sid, n_samples=len(pcm), mode=mode, param=param, avg_complexity=avgc,
            block_size=bs, n_chan=2, rate_hz=RATE,
            data=enc.data, block_bytes=enc.block_bytes, window_ctrl=enc.window_ctrl, complexity=enc.complexity,
            bits=bits,
            dec_sha256=hashlib.sha256(dec.tobytes()).hexdigest(),
            dec_head=dec[:4 * bs].copy(), dec_tail=dec[-2 * bs:].copy(),
            coefs_sha256=hashlib.sha256(enc.coefs.tobytes()).hexdigest(),
            coefs_head=enc.coefs[:6].copy(),
            ranks_sha256=hashlib.sha256(enc.ranks.tobytes()).hexdigest(),
        )
        print(f"{name}: {len(enc.block_bytes)} blocks, {len(enc.data)} bytes, avg {enc.block_bytes.mean():.1f} B/block, "
              f"window codes {sorted(set((enc.window_ctrl >> 4).tolist()))}")


if __name__ == "__main__":
    main()
