"""The drop-in proof (-m gpu): the reference's OWN command-line tools, compiled unmodified from /root/reference/tools by
oracle/Makefile (`make dropin`) with libulc + libfourier replaced by libulcb200.so on the link line, run next to the stock
tools on the same files.  Encoded files must be identical byte for byte (the goldens' criterion), decoded WAVs within 1 LSB."""
import os
import subprocess
import time

import numpy as np
import pytest

from ulc_codec_b200 import corpus, wav

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
TOOLS = {k: os.path.join(REF, k) for k in ("ulcencodetool_scalar", "ulcdecodetool_scalar", "ulcencodetool_b200", "ulcdecodetool_b200")}
needs_tools = pytest.mark.skipif(not all(os.access(p, os.X_OK) for p in TOOLS.values()), reason="oracle/_ref tools not built")


def _run(tool, *a):
    t0 = time.perf_counter()
    r = subprocess.run([TOOLS[tool], *a], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return time.perf_counter() - t0, r.stdout


@needs_tools
@pytest.mark.parametrize("kind,args,seconds", [("music", ["-50"], 12), ("transient", ["-70", "-blocksize:1024"], 12), ("music", ["96"], 12),
                                               ("music", ["64,0.7", "-blocksize:4096"], 12), ("music", ["-50"], 180)])
def test_reference_tools_on_the_gpu_library(tmp_path, kind, args, seconds):
    pcm = corpus.generate(4242, seconds * 44100, kind)
    src = str(tmp_path / "in.wav")
    wav.write(src, wav.WavFormat(wav.WAVE_FORMAT_PCM, 2, 44100, 16), pcm)
    ulc_ref, ulc_gpu = str(tmp_path / "ref.ulc"), str(tmp_path / "gpu.ulc")
    t_ref, _ = _run("ulcencodetool_scalar", src, ulc_ref, *args)
    t_gpu, out = _run("ulcencodetool_b200", src, ulc_gpu, *args)
    a, b = open(ulc_ref, "rb").read(), open(ulc_gpu, "rb").read()
    if seconds <= 30:
        assert a == b, f"encoded files differ ({len(a)} vs {len(b)} bytes)\n{out}"
    else:
        # long streams: BASELINE.json allows a quantisation-boundary flip in at most 0.1 % of blocks (device log vs glibc logf
        # in the last ulp; the reference's own two builds differ in 0.4-1 %)
        from oracle import oracle
        from ulc_codec_b200 import container
        fa, fb = container.parse(a), container.parse(b)
        assert fa.header.n_blocks == fb.header.n_blocks and fa.header.block_size == fb.header.block_size
        nb, bsz = fa.header.n_blocks, fa.header.block_size
        pa, pb = fa.payload, fb.payload
        _, bits_a = oracle.Decoder(2, bsz).decode_stream(pa, nb)
        _, bits_b = oracle.Decoder(2, bsz).decode_stream(pb, nb)
        assert (bits_a != 0).all() and (bits_b != 0).all()
        oa = np.concatenate([[0], np.cumsum((bits_a.astype(np.int64) + 7) // 8)]); ob = np.concatenate([[0], np.cumsum((bits_b.astype(np.int64) + 7) // 8)])
        differing = sum(1 for k in range(nb) if not np.array_equal(pa[oa[k]:oa[k + 1]], pb[ob[k]:ob[k + 1]]))
        assert differing <= max(1, nb // 1000), f"{differing} of {nb} blocks differ"
    wav_ref, wav_gpu = str(tmp_path / "ref.wav"), str(tmp_path / "gpu.wav")
    d_ref, _ = _run("ulcdecodetool_scalar", ulc_ref, wav_ref)
    d_gpu, _ = _run("ulcdecodetool_b200", ulc_ref, wav_gpu)
    wr, wg = wav.read(wav_ref), wav.read(wav_gpu)
    assert wr.n_sample_points == wg.n_sample_points and wr.fmt == wg.fmt
    diff = np.abs(wr.samples().astype(np.int32) - wg.samples().astype(np.int32)).max()
    assert diff <= 1, diff
    print(f"\n[dropin {kind} {' '.join(args)}] encode: stock {seconds / t_ref:.0f} x realtime, libulcb200 {seconds / t_gpu:.0f} x; "
          f"decode: stock {seconds / d_ref:.0f} x, libulcb200 {seconds / d_gpu:.0f} x (single {seconds} s stream, one block per call, process and CUDA start-up included)")
