"""CPU: pin the oracle restatement against the compiled UNMODIFIED reference (oracle/_ref).
Skipped where oracle/_ref is absent; the golden fixtures cover that case."""
import numpy as np
import pytest

from oracle import oracle, refbin
from ulc_codec_b200 import corpus

pytestmark = pytest.mark.skipif(not refbin.available("scalar"), reason="oracle/_ref not built")


@pytest.mark.parametrize("n", [8, 16, 64, 512, 2048, 8192])
def test_dct_bit_identical(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    assert np.array_equal(oracle.dct4(x), refbin.ref_dct4(x))
    assert np.array_equal(oracle.dct2(x), refbin.ref_dct2(x))
    # and both agree with the closed form up to the reference's own twiddle-recurrence
    # error, which grows with n (measured: 6e-6 of peak at 2048, 4.6e-5 at 8192 for the scalar build)
    k = np.arange(n) + 0.5
    exact = np.cos(np.pi / n * np.outer(k, k)) @ x.astype(np.float64)
    assert np.abs(oracle.dct4(x) - exact).max() < (1e-5 if n <= 2048 else 1e-4) * np.abs(exact).max()


@pytest.mark.parametrize("n,ov", [(16, 16), (64, 32), (256, 0), (2048, 2048), (2048, 256), (1024, 64), (8192, 8192)])
def test_mdct_imdct_bit_identical(n, ov):
    rng = np.random.default_rng(n + ov)
    x = rng.standard_normal(n).astype(np.float32)
    lap = rng.standard_normal(n // 2).astype(np.float32)
    l1, l2 = lap.copy(), lap.copy()
    assert np.array_equal(oracle.imdct(x, l1, ov), refbin.ref_imdct(x, l2, ov))
    assert np.array_equal(l1, l2)
    lap = rng.standard_normal(n).astype(np.float32)
    l1, l2 = lap.copy(), lap.copy()
    a, b = oracle.mdct_mdst(x, l1, ov), refbin.ref_mdct_mdst(x, l2, ov)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(l1, l2)


@pytest.mark.parametrize("kind,bs,mode,param", [("music", 2048, "cbr", 128.0), ("transient", 2048, "vbr", 50.0),
                                                ("transient", 256, "vbr", 70.0), ("music", 8192, "vbr", 30.0)])
def test_decoder_bit_identical(kind, bs, mode, param):
    pcm = corpus.generate(11, 44100 * 4, kind)
    enc = refbin.encode(pcm, mode, param, block_size=bs)
    ref, rbits = refbin.decode(enc.data, len(enc.block_bytes), block_size=bs)
    out, bits = oracle.Decoder(2, bs).decode_stream(enc.data, len(enc.block_bytes))
    assert np.array_equal(bits, rbits)
    assert np.array_equal(out, ref)


def test_decoder_mono_and_corrupt():
    pcm = corpus.generate(12, 44100 * 2, "music")[:, :1]
    enc = refbin.encode(pcm, "vbr", 60.0)
    ref, rbits = refbin.decode(enc.data, len(enc.block_bytes), n_chan=1)
    out, bits = oracle.Decoder(1, 2048).decode_stream(enc.data, len(enc.block_bytes))
    assert np.array_equal(out, ref) and np.array_equal(bits, rbits)
    # a run that overshoots the sub-block must be reported as corrupt (ulcDecoder.c:127,139,154)
    bad = np.array([0x00, 0x81, 0xFF, 0xFF] + [0xFF] * 64, np.uint8)   # header 0, quant 0, then 8h,Fh,Fh,Fh = 527 noise... x8
    bad = np.concatenate([np.array([0x00], np.uint8), np.full(64, 0xF8, np.uint8)])
    ref, rbits = refbin.decode(bad, 1, n_chan=1)
    out, bits = oracle.Decoder(1, 2048).decode_stream(bad, 1)
    assert rbits[0] == bits[0]


@pytest.mark.parametrize("kind,bs,mode,param,avgc", [
    ("music", 2048, "vbr", 50.0, 0.0), ("music", 2048, "cbr", 128.0, 0.0), ("transient", 2048, "vbr", 75.0, 0.0),
    ("transient", 256, "cbr", 64.0, 0.0), ("transient", 8192, "vbr", 40.0, 0.0), ("music", 512, "abr", 96.0, 0.8)])
def test_encoder_bit_identical(kind, bs, mode, param, avgc):
    pcm = corpus.generate(21, 44100 * 3, kind)
    ref = refbin.encode(pcm, mode, param, avg_complexity=avgc, block_size=bs, dump=True)
    o = oracle.Encoder(44100, 2, bs).encode_stream(pcm, mode, param, avgc, dump=True)
    assert np.array_equal(o["data"], ref.data)
    assert np.array_equal(o["block_bytes"], ref.block_bytes)
    assert np.array_equal(o["window_ctrl"], ref.window_ctrl)
    assert np.array_equal(o["complexity"], ref.complexity)
    assert np.array_equal(o["coefs"], ref.coefs)
    assert np.array_equal(o["ranks"], ref.ranks)


def test_encoder_mono_and_silence():
    pcm = corpus.generate(22, 30000, "music")[:, :1]
    pcm[8000:16000] = 0                              # digital silence: -inf keys, silent sub-blocks (Eh,Fh)
    ref = refbin.encode(pcm, "vbr", 70.0, dump=True)
    o = oracle.Encoder(44100, 1, 2048).encode_stream(pcm, "vbr", 70.0, dump=True)
    assert np.array_equal(o["data"], ref.data) and np.array_equal(o["ranks"], ref.ranks)


@pytest.mark.parametrize("rate_hz,bs,mode,param", [(32000, 1024, "vbr", 60.0), (48000, 2048, "cbr", 112.0), (22050, 512, "vbr", 45.0), (96000, 4096, "vbr", 70.0)])
def test_encoder_other_sample_rates(rate_hz, bs, mode, param):
    """RateHz enters the Bark tables, the detector filters and the kbps budget."""
    pcm = corpus.generate(31, 9 * bs, "transient" if bs == 1024 else "music")
    ref = refbin.encode(pcm, mode, param, block_size=bs, rate_hz=rate_hz)
    o = oracle.Encoder(rate_hz, 2, bs).encode_stream(pcm, mode, param, 0.0)
    assert np.array_equal(o["data"], ref.data) and np.array_equal(o["window_ctrl"], ref.window_ctrl)


@pytest.mark.parametrize("n_chan", [3, 4, 5])
def test_more_than_two_channels(n_chan):
    """channels are M/S-paired (0,1), (2,3), ...; an odd last channel stays unpaired (ulcEncoder_BlockTransform.c:102).
    The CUDA path does not build these shapes yet; the oracle is pinned for when it does."""
    planes = [corpus.generate(41 + i, 20000, "transient" if i & 1 else "music") for i in range(3)]
    pcm = np.concatenate(planes, axis=1)[:, :n_chan].copy()
    ref = refbin.encode(pcm, "vbr", 60.0, block_size=1024)
    o = oracle.Encoder(44100, n_chan, 1024).encode_stream(pcm, "vbr", 60.0, 0.0)
    assert np.array_equal(o["data"], ref.data)
    d, bits = refbin.decode(ref.data, len(ref.block_bytes), n_chan=n_chan, block_size=1024)
    od, ob = oracle.Decoder(n_chan, 1024).decode_stream(ref.data, len(ref.block_bytes))
    assert np.array_equal(d, od) and np.array_equal(bits, ob)


@pytest.mark.parametrize("bs,n_chan", [(16384, 1), (16384, 2), (32768, 1)])
def test_largest_block_sizes(bs, n_chan):
    """BlockSize 16384 / 32768 (ulcEncoder.c:18-34, ulcDecoder.c:30-35).  The CUDA path builds 16384 in mono; the oracle
    is pinned for all of them."""
    pcm = corpus.generate(51, 5 * bs + 4321, "music")[:, :n_chan].copy()
    ref = refbin.encode(pcm, "vbr", 50.0, block_size=bs)
    o = oracle.Encoder(44100, n_chan, bs).encode_stream(pcm, "vbr", 50.0, 0.0)
    assert np.array_equal(o["data"], ref.data) and np.array_equal(o["window_ctrl"], ref.window_ctrl)
    d, bits = refbin.decode(ref.data, len(ref.block_bytes), n_chan=n_chan, block_size=bs)
    od, ob = oracle.Decoder(n_chan, bs).decode_stream(ref.data, len(ref.block_bytes))
    assert np.array_equal(d, od) and np.array_equal(bits, ob)
