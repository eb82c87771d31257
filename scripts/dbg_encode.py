"""GPU debug: encode golden streams with taps and compare every stage with the oracle."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from tests.fixtures import load_golden_stream, golden_pcm, GOLDEN_NAMES
from ulc_codec_b200 import capi

names = sys.argv[1:] or GOLDEN_NAMES
for name in names:
    g = load_golden_stream(name)
    N, C_ = g["block_size"], g["n_chan"]
    pcm = golden_pcm(g)
    o = oracle.Encoder(g["rate_hz"], C_, N).encode_stream(pcm, g["mode"], g["param"], g["avg_complexity"], dump="all")
    nb = len(o["block_bytes"])
    with capi.Context(g["rate_hz"], C_, N, max_blocks=nb) as ctx:
        r = ctx.encode_streams([pcm], g["mode"], g["param"], g["avg_complexity"], taps=True)
    def cmp(label, a, b, per_block=True):
        a = np.asarray(a); b = np.asarray(b)
        if a.shape != b.shape:
            print(f"  {label}: SHAPE {a.shape} vs {b.shape}"); return
        same = (a == b) | ((a != a) & (b != b))
        if same.all():
            print(f"  {label}: identical"); return
        if per_block:
            bad = np.nonzero(~same.reshape(len(a), -1).all(axis=1))[0]
            fb = bad[0]
            idx = np.nonzero(~same[fb].reshape(-1))[0]
            print(f"  {label}: {len(bad)}/{len(a)} blocks differ, first block {fb} at {idx[:8]} got {a[fb].reshape(-1)[idx[:4]]} want {b[fb].reshape(-1)[idx[:4]]}")
        else:
            print(f"  {label}: differ")
    print(name, "blocks", nb)
    cmp("window_ctrl", r["window_ctrl"][0], o["window_ctrl"], False)
    if not np.array_equal(r["window_ctrl"][0], o["window_ctrl"]):
        bad = np.nonzero(r["window_ctrl"][0] != o["window_ctrl"])[0]
        print("     first wc diffs", [(int(b), hex(r["window_ctrl"][0][b]), hex(o["window_ctrl"][b])) for b in bad[:6]])
    cmp("coefs", r["coefs"][0], o["coefs"])
    cmp("complexity", r["complexity"][0], o["complexity"], False)
    if not np.array_equal(r["complexity"][0], o["complexity"]):
        bad = np.nonzero(r["complexity"][0] != o["complexity"])[0]
        print("     cx diffs", len(bad), [(int(b), float(r["complexity"][0][b]), float(o["complexity"][b])) for b in bad[:4]])
    cmp("noise", r["noise"][0], o["noise"])
    cmp("mask", r["mask"][0], o["mask"])
    cmp("keys", r["keys"][0], o["keys"])
    cmp("n_nz", r["counts"][0][:, 0], o["n_nz"], False)
    cmp("n_out", r["counts"][0][:, 1], o["n_out"], False)
    # ranks: compare only among codeable coefficients
    rk_bad = 0
    for b in range(nb):
        nz = o["n_nz"][b]
        m = o["ranks"][b] < nz
        if not np.array_equal(r["ranks"][0][b][m], o["ranks"][b][m]): rk_bad += 1
    print(f"  ranks (codeable only): {rk_bad}/{nb} blocks differ")
    cmp("block_bytes", r["block_bytes"][0], o["block_bytes"], False)
    same_bytes = np.array_equal(r["streams"][0], o["data"])
    if not same_bytes:
        offs_o = np.concatenate([[0], np.cumsum(o["block_bytes"])]); offs_g = np.concatenate([[0], np.cumsum(r["block_bytes"][0])])
        nbad = 0; first = None
        for b in range(nb):
            if not np.array_equal(r["streams"][0][offs_g[b]:offs_g[b+1]], o["data"][offs_o[b]:offs_o[b+1]]):
                nbad += 1
                if first is None: first = b
        print(f"  bytes: {nbad}/{nb} blocks differ, first {first}")
        b = first
        gg = r["streams"][0][offs_g[b]:offs_g[b+1]]; oo = o["data"][offs_o[b]:offs_o[b+1]]
        n = min(len(gg), len(oo)); d = np.nonzero(gg[:n] != oo[:n])[0]
        print("     first differing byte", d[:5], "len", len(gg), len(oo), "got", gg[d[0]-2:d[0]+6] if len(d) else None, "want", oo[d[0]-2:d[0]+6] if len(d) else None)
    else:
        print("  bytes: identical")
