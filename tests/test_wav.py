"""RIFF/WAVE framing: ulc_codec_b200/wav.py against the UNMODIFIED reference tools (tools/WavIO_Reader.c,
tools/WavIO_Writer.c, tools/MiniRIFF.c as compiled into oracle/_ref/ulc{en,de}codetool_scalar).  CPU only."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests.fixtures import load_golden_stream
from ulc_codec_b200 import container, wav

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
ENC_TOOL, DEC_TOOL = os.path.join(REF, "ulcencodetool_scalar"), os.path.join(REF, "ulcdecodetool_scalar")
needs_tools = pytest.mark.skipif(not (os.access(ENC_TOOL, os.X_OK) and os.access(DEC_TOOL, os.X_OK)), reason="oracle/_ref tools not built")
FORMATS = {"PCM8": (wav.WAVE_FORMAT_PCM, 8), "PCM16": (wav.WAVE_FORMAT_PCM, 16), "PCM24": (wav.WAVE_FORMAT_PCM, 24), "FLOAT32": (wav.WAVE_FORMAT_IEEE_FLOAT, 32)}


@needs_tools
@pytest.mark.parametrize("name", list(FORMATS))
def test_reader_and_writer_agree_with_the_reference_decoder_tool(tmp_path, name):
    """The reference writer's file parses to the right format and sample count, and build() reproduces it byte for byte."""
    g = load_golden_stream("music_vbr75_bs2048")
    nb = 9
    offs = np.concatenate([[0], np.cumsum(g["block_bytes"])]).astype(np.int64)
    ulc, out = str(tmp_path / "a.ulc"), str(tmp_path / "a.wav")
    container.write(ulc, 2048, 44100, 2, g["data"][:offs[nb]], g["block_bytes"][:nb])
    subprocess.run([DEC_TOOL, ulc, out, "-format:" + name], check=True, stdout=subprocess.DEVNULL)
    file_bytes = open(out, "rb").read()
    w = wav.parse(file_bytes)
    tag, bits = FORMATS[name]
    assert (w.fmt.tag, w.fmt.bits, w.fmt.n_channels, w.fmt.rate_hz) == (tag, bits, 2, 44100)
    assert w.n_sample_points == (nb - 1) * 2048 or w.n_sample_points == nb * 2048      # the tool drops the coding delay
    assert w.samples().shape[:2] == (w.n_sample_points, 2)
    assert wav.build(w.fmt, w.raw) == file_bytes


@needs_tools
def test_reference_encoder_tool_reads_our_files(tmp_path):
    """A file written here (with an odd-sized extra chunk in front of the data and 8-bit mono padding) reads like a plain one."""
    from ulc_codec_b200 import corpus
    pcm = corpus.generate(77, 2048 * 6, "music")
    a, b = str(tmp_path / "a.wav"), str(tmp_path / "b.wav")
    fmt = wav.WavFormat(wav.WAVE_FORMAT_PCM, 2, 44100, 16)
    wav.write(a, fmt, pcm)
    plain = open(a, "rb").read()
    # same payload behind a LIST chunk of odd size (padded to 2 bytes, tools/MiniRIFF.c:15-17)
    extra = b"LIST" + struct.pack("<I", 7) + b"INFOabc" + b"\0"
    body = plain[8:36] + extra + plain[36:]
    open(b, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    outs = []
    for src in (a, b):
        subprocess.run([ENC_TOOL, src, src + ".ulc", "-60", "-blocksize:1024"], check=True, stdout=subprocess.DEVNULL)
        outs.append(open(src + ".ulc", "rb").read())
    assert outs[0] == outs[1]
    w = wav.read(b)
    assert w.n_sample_points == len(pcm) and np.array_equal(w.samples(), pcm)


def test_sample_count_rules_and_errors():
    fmt = wav.WavFormat(wav.WAVE_FORMAT_PCM, 3, 48000, 24)
    raw = np.arange(3 * 3 * 5 + 4, dtype=np.uint8)                   # 5 whole sample points and 4 stray bytes
    w = wav.parse(wav.build(fmt, raw))
    assert w.n_sample_points == 5 and w.samples().shape == (5, 3, 3)
    odd = wav.build(wav.WavFormat(wav.WAVE_FORMAT_PCM, 1, 8000, 8), np.arange(7, dtype=np.uint8))
    assert len(odd) % 2 == 0 and wav.parse(odd).n_sample_points == 7  # data chunk padded to 2 bytes (WavIO_Helper.c:131)
    with pytest.raises(wav.WavError):
        wav.parse(b"RIFF\x04\0\0\0WAVX")
    with pytest.raises(wav.WavError):
        wav.parse(wav.build(fmt, raw).replace(b"data", b"dat_"))
    with pytest.raises(wav.WavError):
        wav.build(wav.WavFormat(wav.WAVE_FORMAT_PCM, 2, 44100, 32), b"")    # PCM32 is not a format the reference reads
    zero_ch = bytearray(wav.build(wav.WavFormat(wav.WAVE_FORMAT_PCM, 1, 8000, 16), b"\0\0"))
    zero_ch[22:24] = b"\0\0"
    with pytest.raises(wav.WavError):
        wav.parse(bytes(zero_ch))
