"""Sample-format conversion: the oracle's numpy restatement of tools/WavIO_Helper.c:31-87 is pinned
against the UNMODIFIED reference tools (CPU), and the CUDA kernels behind ULCB_ConvertSamplesDevice are held
bit-exact to it (GPU)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import oracle
from tests.fixtures import load_golden_stream

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
ENC_TOOL, DEC_TOOL = os.path.join(REF, "ulcencodetool_scalar"), os.path.join(REF, "ulcdecodetool_scalar")
needs_tools = pytest.mark.skipif(not (os.access(ENC_TOOL, os.X_OK) and os.access(DEC_TOOL, os.X_OK)), reason="oracle/_ref tools not built")
FORMATS = {"pcm8u": ("PCM8", 1, 8, 1), "pcm16": ("PCM16", 1, 16, 2), "pcm24": ("PCM24", 1, 24, 3), "f32": ("FLOAT32", 3, 32, 4)}


def _wav(raw: bytes, tag: int, bits: int, nch: int, rate: int = 44100) -> bytes:
    bps = bits // 8
    return (b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, tag, nch, rate, rate * nch * bps, nch * bps, bits)
            + b"data" + struct.pack("<I", len(raw)) + raw)


def _wav_data(path: str) -> bytes:
    raw = open(path, "rb").read()
    i = raw.index(b"data")
    return raw[i + 8:i + 8 + struct.unpack_from("<I", raw, i + 4)[0]]


def _edge_values() -> np.ndarray:
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1.2, 1.2, 4000), rng.normal(0, 1e-3, 1000),
                        (np.arange(-300, 300) + 0.5) / 128.0, (np.arange(-300, 300) + 0.5) / 32768.0, (np.arange(-300, 300) + 0.5) / 8388608.0,
                        [0.0, -0.0, 1.0, -1.0, 0.99999994, 127.5 / 128, 32767.5 / 32768, 2.0, -2.0]]).astype(np.float32)
    return x[: len(x) // 2 * 2]


@needs_tools
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_from_float_matches_reference_decoder_tool(tmp_path, fmt):
    """ulcdecodetool -format:X writes WAV_ConvertFromFloat_X(decoded block); the restatement must give the same bytes."""
    g = load_golden_stream("music_vbr75_bs2048")
    nb = 20
    offs = np.concatenate([[0], np.cumsum(g["block_bytes"])]).astype(np.int64)
    from ulc_codec_b200 import container
    ulc, wav = str(tmp_path / "a.ulc"), str(tmp_path / "a.wav")
    container.write(ulc, 2048, 44100, 2, g["data"][:offs[nb]], g["block_bytes"][:nb])
    subprocess.run([DEC_TOOL, ulc, wav, "-format:" + FORMATS[fmt][0]], check=True, stdout=subprocess.DEVNULL)
    ref, _ = oracle.Decoder(2, 2048).decode_stream(g["data"], nb)
    want = oracle.samples_from_float(ref.reshape(-1), fmt)
    assert _wav_data(wav) == want.tobytes()


@needs_tools
@pytest.mark.parametrize("fmt", ["pcm8u", "pcm24", "f32"])
def test_to_float_matches_reference_encoder_tool(tmp_path, fmt):
    """ulcencodetool reads a WAV in format X through WAV_ConvertToFloat_X; encoding the restatement's floats (as a
    float WAV, an identity conversion) must give the same file."""
    from ulc_codec_b200 import corpus
    pcm = corpus.generate(31, 12000, "music").astype(np.float32) / np.float32(32768.0)
    raw = oracle.samples_from_float(pcm.reshape(-1), fmt)
    _, tag, bits, _ = FORMATS[fmt]
    a_wav, b_wav, a_ulc, b_ulc = (str(tmp_path / n) for n in ("a.wav", "b.wav", "a.ulc", "b.ulc"))
    open(a_wav, "wb").write(_wav(raw.tobytes(), tag, bits, 2))
    open(b_wav, "wb").write(_wav(oracle.samples_to_float(raw, fmt).tobytes(), 3, 32, 2))
    for w, u in ((a_wav, a_ulc), (b_wav, b_ulc)):
        subprocess.run([ENC_TOOL, w, u, "-60", "-blocksize:512"], check=True, stdout=subprocess.DEVNULL)
    assert open(a_ulc, "rb").read() == open(b_ulc, "rb").read()


def test_restatement_roundtrips():
    x = _edge_values()
    for fmt, (_, _, bits, _) in FORMATS.items():
        raw = oracle.samples_from_float(x, fmt)
        back = oracle.samples_to_float(raw, fmt)
        if fmt == "f32":
            assert np.array_equal(back, x)
            continue
        if fmt == "pcm8u":                                                           # the reference's 8-bit reader does not invert its writer
            assert np.array_equal(oracle.samples_to_float(np.array([0x00, 0x7F, 0x80, 0xFF], np.uint8), fmt),
                                  np.array([128, 255, -256, -129], np.float32) / 128)
            continue
        step = 2.0 ** -(bits - 1)
        inside = np.abs(x) < 1.0 - step
        assert np.abs(back[inside] - x[inside]).max() <= step / 2 + 1e-9
        assert np.array_equal(oracle.samples_from_float(back, fmt), raw)           # idempotent


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_gpu_conversion_is_bit_exact(fmt):
    import torch
    from ulc_codec_b200 import capi
    dev = torch.device("cuda", 0)
    x = _edge_values()
    for n in (len(x), 4093, 7, 1):                                                  # ragged tails, odd alignments
        xs = x[:n]
        want_raw = oracle.samples_from_float(xs, fmt)
        d_x = torch.from_numpy(xs.copy()).to(dev)
        nbytes = len(want_raw.tobytes())
        d_raw = torch.zeros(nbytes + 16, dtype=torch.uint8, device=dev)
        d_back = torch.zeros(n, dtype=torch.float32, device=dev)
        with capi.Context(44100, 2, 2048, max_blocks=4) as ctx:
            ctx.convert_samples_device(fmt, False, d_x.data_ptr(), d_raw.data_ptr(), n)
            ctx.convert_samples_device(fmt, True, d_raw.data_ptr(), d_back.data_ptr(), n)
            ctx.sync()
        got_raw = d_raw.cpu().numpy()
        assert got_raw[:nbytes].tobytes() == want_raw.tobytes(), (fmt, n)
        assert not got_raw[nbytes:].any()                                           # nothing written past the end
        assert np.array_equal(d_back.cpu().numpy(), oracle.samples_to_float(want_raw, fmt)), (fmt, n)
