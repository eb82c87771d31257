import sys, numpy as np
sys.path.insert(0, '.')
from ulc_codec_b200 import capi
from oracle import oracle
from tests.fixtures import load_golden_stream
name = sys.argv[1] if len(sys.argv) > 1 else "music_cbr128_bs2048"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 12
g = load_golden_stream(name); N = g["block_size"]; C = g["n_chan"]
d = oracle.Decoder(C, N)
buf = np.concatenate([g["data"], np.zeros(70000, np.uint8)])
pos = 0; ref = []; rbits = []
for b in range(nb):
    c, wc, bits = d.parse_block(buf, pos); ref.append(c); rbits.append(bits); pos += (bits + 7)//8
ref = np.stack(ref)
with capi.Context(g["rate_hz"], C, N, max_blocks=nb) as ctx:
    coefs, bits = ctx.test_parse(g["data"], nb)
offs = np.concatenate([[0], np.cumsum(g["block_bytes"])])
for b in range(nb):
    bad = np.nonzero(coefs[b] != ref[b])[0]
    print(b, "wc=%02x" % g["window_ctrl"][b], "byte", int(offs[b]), "bits", bits[b], rbits[b], "nbad", len(bad), "first", bad[:6], "nz gpu/ref", np.count_nonzero(coefs[b]), np.count_nonzero(ref[b]))
    if len(bad):
        i = bad[0]; print("   gpu", coefs[b, i:i+8], "\n   ref", ref[b, i:i+8])
