import sys, numpy as np
sys.path.insert(0, '.')
from ulc_codec_b200 import capi
from oracle import oracle
from tests.fixtures import load_golden_stream, peak_rel_err
name = sys.argv[1] if len(sys.argv) > 1 else "music_cbr128_bs2048"
g = load_golden_stream(name)
nb = len(g["block_bytes"]); N = g["block_size"]
ref, rbits = oracle.Decoder(g["n_chan"], N).decode_stream(g["data"], nb)
with capi.Context(g["rate_hz"], g["n_chan"], N, max_blocks=nb) as ctx:
    pcm, status, bits = ctx.decode_streams([g["data"]], nb, fmt="f32", want_bits=True)
print("status", status, "first bits mismatch", np.nonzero(bits[0] != rbits)[0][:5])
offs = np.concatenate([[0], np.cumsum(g["block_bytes"])])
peak = np.abs(ref).max()
for b in range(nb):
    e = np.abs(pcm[0, b*N:(b+1)*N] - ref[b*N:(b+1)*N]).max() / peak
    flag = "" if e < 1e-5 else "  <<<<"
    if b < 8 or flag or bits[0, b] != rbits[b]:
        print(b, "wc=%02x" % g["window_ctrl"][b], "off", offs[b], "page", (offs[b]*2)//2048, "bits", bits[0, b], rbits[b], "err %.3g" % e, flag)
    if b > 60 and flag: break
