"""Two-pass ABR orchestration: ULCB_MODE_ABR_2PASS measures each stream's AvgComplexity in an analysis
pass on the device and feeds it to the ABR pass - what a user of the reference tool does by hand (encode once, read
"Avg complexity = %.5f", encode again with RateKbps,AvgComplexity; tools/ulcEncodeTool.c:41,164,177-188)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from oracle import oracle
from ulc_codec_b200 import corpus

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
ENC_TOOL = os.path.join(REF, "ulcencodetool_scalar")


def tool_avg_complexity(cx: np.ndarray) -> float:
    """What the second run of the tool receives: mean of the block complexities in double (:164,:177), printed with
    %.5f (:187), typed back and read with %f (:41)."""
    s = 0.0
    for v in cx.astype(np.float32):
        s += float(v)
    return float(np.float32(float("%.5f" % (s / len(cx)))))


def _wav(pcm16, rate=44100):
    data = np.ascontiguousarray(pcm16, np.int16).tobytes()
    nch = pcm16.shape[1]
    return (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, nch, rate, rate * nch * 2, nch * 2, 16)
            + b"data" + struct.pack("<I", len(data)) + data)


@pytest.mark.skipif(not os.access(ENC_TOOL, os.X_OK), reason="oracle/_ref tools not built")
def test_orchestration_matches_the_reference_tool(tmp_path):
    """Pins tool_avg_complexity() and the oracle's two-pass result against two runs of the unmodified tool."""
    pcm = corpus.generate(91, 40000, "transient")
    wav, u1, u2 = (str(tmp_path / n) for n in ("a.wav", "p1.ulc", "p2.ulc"))
    open(wav, "wb").write(_wav(pcm))
    out = subprocess.run([ENC_TOOL, wav, u1, "72", "-blocksize:1024"], check=True, capture_output=True, text=True).stdout
    printed = re.search(r"Avg complexity = ([0-9.]+)", out).group(1)
    p1 = oracle.Encoder(44100, 2, 1024).encode_stream(pcm, "cbr", 72.0)
    assert "%.5f" % float(printed) == printed and tool_avg_complexity(p1["complexity"]) == float(np.float32(float(printed)))
    subprocess.run([ENC_TOOL, wav, u2, "72," + printed, "-blocksize:1024"], check=True, stdout=subprocess.DEVNULL)
    p2 = oracle.Encoder(44100, 2, 1024).encode_stream(pcm, "abr", 72.0, tool_avg_complexity(p1["complexity"]))
    assert open(u2, "rb").read()[24:] == p2["data"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("bs,kbps", [(2048, 64.0), (512, 96.0)])
def test_gpu_two_pass_abr(bs, kbps):
    from ulc_codec_b200 import capi
    pcms = [corpus.generate(300 + i, 30000 + 4000 * i, "transient" if i % 2 else "music") for i in range(5)]
    with capi.Context(44100, 2, bs, max_blocks=64) as ctx:
        nb = ctx.blocks_for(max(len(p) for p in pcms))
        g = ctx.encode_streams(pcms, "abr2", kbps, n_blocks=nb)
        avg = ctx.avg_complexity(len(pcms))
        first = ctx.encode_streams(pcms, "cbr", kbps, n_blocks=nb)              # complexity does not depend on the rate mode
    assert not g["status"].any()
    for i, pcm in enumerate(pcms):
        x = np.zeros(((nb - 2) * bs, 2), np.int16); x[:len(pcm)] = pcm           # same block count as the batch
        p1 = oracle.Encoder(44100, 2, bs).encode_stream(x, "cbr", kbps)
        want_avg = tool_avg_complexity(p1["complexity"])
        assert np.array_equal(first["complexity"][i], p1["complexity"])
        assert avg[i] == np.float32(want_avg), (i, avg[i], want_avg)
        p2 = oracle.Encoder(44100, 2, bs).encode_stream(x, "abr", kbps, want_avg)
        assert np.array_equal(g["block_bytes"][i], p2["block_bytes"]), i
        assert g["streams"][i].tobytes() == p2["data"].tobytes(), i
