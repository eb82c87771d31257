"""GPU fuzz: corrupted streams must neither hang the decoder nor disagree with the oracle about which block is the first
undecodable one, and every block before the first corrupted byte must decode the same.  (A corrupted block that still
parses decodes to garbage in the reference too; that garbage is not part of the contract and is not compared.)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle
from ulc_codec_b200 import capi
from tests.fixtures import load_golden_stream

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
nb = 24
streams, refs, first_hit = [], [], []
for i in range(96):
    g = load_golden_stream(["music_cbr128_bs2048", "transient_vbr50_bs2048", "music_vbr75_bs2048"][i % 3])
    offs = np.concatenate([[0], np.cumsum(g["block_bytes"])]).astype(np.int64)
    s = g["data"][:offs[nb]].copy()
    kind = i % 4
    orig = s.copy()
    if kind == 0:
        for _ in range(1 + i % 5): s[rng.integers(0, len(s))] = rng.integers(0, 256)
    elif kind == 1:
        a = int(rng.integers(0, len(s) - 64)); s[a:a + int(rng.integers(1, 64))] = 0xFF
    elif kind == 2:
        s = s[:int(rng.integers(1, len(s)))]                      # truncated
    else:
        a = int(rng.integers(0, len(s) - 64)); s[a:a + int(rng.integers(1, 64))] = 0x88
    streams.append(s)
    m = min(len(s), len(orig)); diff = np.flatnonzero(s[:m] != orig[:m])
    hit = int(diff[0]) if len(diff) else (len(s) if len(s) < len(orig) else 1 << 60)
    first_hit.append(int(np.searchsorted(offs, hit, side='right') - 1))      # block that holds the first changed byte
    refs.append(oracle.Decoder(2, 2048).decode_stream(np.concatenate([s, np.zeros(64, np.uint8)]), nb))
with capi.Context(44100, 2, 2048, max_blocks=nb) as ctx:
    pcm, status, bits = ctx.decode_streams(streams, nb, fmt="f32", want_bits=True)
bad = 0
for i, (ref, rbits) in enumerate(refs):
    dead = int(status[i])
    first_bad = int(np.flatnonzero(rbits == 0)[0]) if (rbits == 0).any() else -1
    # a truncated stream may end inside a block: the reference reads past the end (zeros appended here), the batched API stops
    good = dead if dead >= 0 else nb
    if first_bad >= 0 and dead >= 0 and dead > first_bad: bad += 1; print("stream", i, "gpu flagged", dead, "oracle", first_bad)
    ok_upto = min(good, first_bad if first_bad >= 0 else nb)
    ends = np.cumsum((rbits[:ok_upto].astype(np.int64) + 7) // 8)
    ok_upto = int((ends <= len(streams[i])).sum())           # only blocks that lie wholly inside a truncated stream
    ok_upto = min(ok_upto, first_hit[i])                      # output block k mixes blocks k-1 and k: compare before the hit
    if ok_upto > 0:
        e = np.abs(pcm[i, :ok_upto * 2048] - ref[:ok_upto * 2048]).max() / max(1e-9, np.abs(ref).max())
        if not np.array_equal(bits[i, :ok_upto], rbits[:ok_upto]) or e > 2e-5:
            bad += 1
            per = np.abs(pcm[i, :ok_upto * 2048] - ref[:ok_upto * 2048]).reshape(ok_upto, -1).max(axis=1)
            fb = int(np.flatnonzero(per > 1e-4 * max(1e-9, np.abs(ref).max()))[0]) if (per > 1e-4 * max(1e-9, np.abs(ref).max())).any() else -1
            print("stream", i, "kind", i % 4, "mismatch before the bad block", e, "first differing block", fb, "of", ok_upto, "dead", dead, "oracle bad", first_bad,
                  "bits equal", np.array_equal(bits[i, :ok_upto], rbits[:ok_upto]), "peak", float(np.abs(ref).max()), "gpu peak", float(np.abs(pcm[i, :ok_upto*2048]).max()))
# Beyond the contract's letter: a damaged block that still parses is decoded by the reference too, and a decoder is a pure function
# of the bytes - so every block both sides accept is compared as well (output block k mixes blocks k-1 and k; the reference's
# RNG position after a damaged block is part of that function).  Mismatching streams are dumped for minimisation.
garbage = 0
for i, (ref, rbits) in enumerate(refs):
    dead = int(status[i]); good = dead if dead >= 0 else nb
    first_bad = int(np.flatnonzero(rbits == 0)[0]) if (rbits == 0).any() else nb
    upto = min(good, first_bad)
    ends = np.cumsum((rbits[:upto].astype(np.int64) + 7) // 8)
    upto = int((ends <= len(streams[i])).sum())
    if upto <= 0: continue
    peak = max(1e-9, float(np.abs(ref[:upto * 2048]).max()))
    per = np.abs(pcm[i, :upto * 2048] - ref[:upto * 2048]).reshape(upto, -1).max(axis=1) / peak
    if not np.array_equal(bits[i, :upto], rbits[:upto]) or (per > 2e-5).any():
        garbage += 1
        fb = int(np.flatnonzero(per > 2e-5)[0]) if (per > 2e-5).any() else -1
        print("stream", i, "kind", i % 4, "parseable-but-damaged block differs: first differing block", fb, "of", upto, "first hit block", first_hit[i],
              "err", float(per.max()), "peak", peak)
        os.makedirs("gpurun_out", exist_ok=True)
        np.savez(f"gpurun_out/fuzz_case_{sys.argv[1] if len(sys.argv) > 1 else 7}_{i}.npz", data=streams[i], nb=nb, first=fb)
print("fuzz done:", len(streams), "streams,", bad, "disagreements;", garbage, "streams with a differing damaged-but-parseable block;", int((status >= 0).sum()), "flagged")
bad += garbage
sys.exit(1 if bad else 0)
