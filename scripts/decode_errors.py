#!/usr/bin/env python
"""Print the decoder's error against the oracle (scalar reference) and, when present, the AVX reference build, per golden
stream: float peak-relative error and PCM16 LSB histogram.  Usage: python scripts/decode_errors.py [out.json]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle, refbin
from tests.fixtures import GOLDEN_NAMES, load_golden_stream, peak_rel_err
from ulc_codec_b200 import capi

res = {}
for name in GOLDEN_NAMES:
    g = load_golden_stream(name)
    nb = len(g["block_bytes"])
    ref, _ = oracle.Decoder(g["n_chan"], g["block_size"]).decode_stream(g["data"], nb)
    with capi.Context(g["rate_hz"], g["n_chan"], g["block_size"], max_blocks=nb) as ctx:
        f, st, _ = ctx.decode_streams([g["data"]], nb, fmt="f32")
        p, _, _ = ctx.decode_streams([g["data"]], nb, fmt="pcm16")
    d = np.abs(p[0].astype(np.int32) - oracle.float_to_pcm16(ref).astype(np.int32))
    r = {"N": g["block_size"], "f32_vs_scalar": peak_rel_err(f[0], ref), "pcm16_max_lsb": int(d.max()), "pcm16_frac_off": float((d > 0).mean()),
         "pcm16_frac_gt1": float((d > 1).mean())}
    if refbin.available("avx"):
        ra, _ = refbin.decode(g["data"], nb, block_size=g["block_size"], variant="avx")
        r["f32_vs_avx"] = peak_rel_err(f[0], ra)
        r["scalar_vs_avx"] = peak_rel_err(ref, ra)
        da = np.abs(p[0].astype(np.int32) - oracle.float_to_pcm16(ra).astype(np.int32))
        r["pcm16_max_lsb_vs_avx"] = int(da.max())
    res[name] = r
    print(name, json.dumps(r))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
