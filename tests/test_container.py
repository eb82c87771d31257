"""The .ulc container: ulc_codec_b200/container.py against files written and read by the UNMODIFIED
reference tools (oracle/_ref/ulcencodetool_scalar, ulcdecodetool_scalar; tools/ulc_Helper.h:10-20)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tests.fixtures import load_golden_stream
from ulc_codec_b200 import container

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
ENC_TOOL, DEC_TOOL = os.path.join(REF, "ulcencodetool_scalar"), os.path.join(REF, "ulcdecodetool_scalar")
needs_tools = pytest.mark.skipif(not (os.access(ENC_TOOL, os.X_OK) and os.access(DEC_TOOL, os.X_OK)), reason="oracle/_ref tools not built")


def _wav(pcm16, rate):
    data = np.ascontiguousarray(pcm16, np.int16).tobytes()
    nch = pcm16.shape[1]
    return (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, nch, rate, rate * nch * 2, nch * 2, 16)
            + b"data" + struct.pack("<I", len(data)) + data)


def _read_wav_pcm16(path):
    raw = open(path, "rb").read()
    i = raw.index(b"data")
    n = struct.unpack_from("<I", raw, i + 4)[0]
    return np.frombuffer(raw, np.int16, count=n // 2, offset=i + 8)


def test_header_roundtrip_and_validation():
    h = container.UlcHeader(block_size=2048, n_blocks=1294, rate_hz=44100, n_chan=2, max_block_size=913, rate_kbps=128)
    raw = h.pack()
    assert len(raw) == 24 and raw[:4] == b"ULC2"
    assert container.UlcHeader.unpack(raw) == h
    assert container.n_blocks_for(60 * 44100, 2048) == 1294                   # tools/ulcEncodeTool.c:98
    with pytest.raises(container.UlcFormatError):
        container.UlcHeader.unpack(b"RIFF" + raw[4:])
    with pytest.raises(container.UlcFormatError):
        container.UlcHeader.unpack(raw[:20])
    bad = bytearray(raw); struct.pack_into("<H", bad, 4, 1000)                 # BlockSize not a power of two
    with pytest.raises(container.UlcFormatError):
        container.UlcHeader.unpack(bytes(bad))
    with pytest.raises(container.UlcFormatError):
        container.UlcHeader(block_size=2048, n_blocks=1, rate_hz=44100, n_chan=2, max_block_size=70000).pack()


def test_block_index_and_batch_packing():
    g = load_golden_stream("music_cbr128_bs2048")
    bb = g["block_bytes"].astype(np.int64)
    offs = container.block_index(bb * 8 - 3)                                   # bits are not byte-rounded (ulcDecoder.c:301)
    assert np.array_equal(offs, np.concatenate([[0], np.cumsum(bb)]))
    assert len(container.block_index([800, 0, 800])) == 2                      # a corrupt block ends the index
    raw = container.build(2048, 44100, 2, g["data"], bb)
    f = container.parse(raw)
    assert f.header.n_blocks == len(bb) and f.header.max_block_size == bb.max() and f.header.stream_offs == 24
    assert np.array_equal(f.payload, g["data"])
    f.offsets = offs
    assert np.array_equal(f.block(3), g["data"][offs[3]:offs[4]])
    data, soffs, nb = container.pack_batch([f, f])
    assert len(data) == 2 * len(g["data"]) and list(soffs) == [0, len(g["data"]), 2 * len(g["data"])] and list(nb) == [len(bb)] * 2
    other = container.parse(container.build(1024, 44100, 2, g["data"], bb))
    with pytest.raises(container.UlcFormatError):
        container.pack_batch([f, other])
    with pytest.raises(container.UlcFormatError):
        container.build(2048, 44100, 2, g["data"][:-1], bb)


@needs_tools
def test_reads_what_the_reference_encoder_writes(tmp_path):
    from ulc_codec_b200 import corpus
    pcm = corpus.generate(77, 30000, "music")
    wav, ulc = str(tmp_path / "a.wav"), str(tmp_path / "a.ulc")
    open(wav, "wb").write(_wav(pcm, 44100))
    subprocess.run([ENC_TOOL, wav, ulc, "96", "-blocksize:1024"], check=True, stdout=subprocess.DEVNULL)
    f = container.read(ulc)
    h = f.header
    assert (h.block_size, h.rate_hz, h.n_chan, h.stream_offs) == (1024, 44100, 2, 24)
    assert h.n_blocks == container.n_blocks_for(30000, 1024)
    # header statistics are what build() derives from the block sizes (the oracle restatement parses the stream)
    from oracle import oracle
    _, bits = oracle.Decoder(2, 1024).decode_stream(f.payload, h.n_blocks)
    offs = container.block_index(bits)
    assert offs[-1] == len(f.payload) and len(offs) == h.n_blocks + 1
    bb = np.diff(offs)
    assert h.max_block_size == bb.max()
    assert container.build(1024, 44100, 2, f.payload, bb) == open(ulc, "rb").read()     # byte-identical file


@needs_tools
def test_reference_decoder_reads_what_we_write(tmp_path):
    """A golden (reference-encoded) block stream wrapped by container.write decodes in the reference tool to the
    PCM16 the oracle predicts."""
    from oracle import oracle
    g = load_golden_stream("music_vbr75_bs2048")
    nb = 40
    offs = np.concatenate([[0], np.cumsum(g["block_bytes"])]).astype(np.int64)
    ulc, wav = str(tmp_path / "b.ulc"), str(tmp_path / "b.wav")
    container.write(ulc, 2048, 44100, 2, g["data"][:offs[nb]], g["block_bytes"][:nb])
    subprocess.run([DEC_TOOL, ulc, wav], check=True, stdout=subprocess.DEVNULL)
    got = _read_wav_pcm16(wav).reshape(-1, 2)
    ref, _ = oracle.Decoder(2, 2048).decode_stream(g["data"], nb)
    assert np.array_equal(got, oracle.float_to_pcm16(ref)[:len(got)]) and len(got) == nb * 2048


@pytest.mark.gpu
def test_gpu_decodes_a_batch_of_ulc_files_and_indexes_them():
    """Files in, PCM + block index out: pack_batch -> ULCB_DecodeStreams (blockBits) -> block_index."""
    from oracle import oracle
    from ulc_codec_b200 import capi
    files, refs = [], []
    for name, nb in (("music_vbr75_bs2048", 30), ("transient_vbr50_bs2048", 45), ("music_cbr128_bs2048", 12)):
        g = load_golden_stream(name)
        offs = np.concatenate([[0], np.cumsum(g["block_bytes"])]).astype(np.int64)
        files.append(container.parse(container.build(2048, 44100, 2, g["data"][:offs[nb]], g["block_bytes"][:nb])))
        refs.append((oracle.float_to_pcm16(oracle.Decoder(2, 2048).decode_stream(g["data"], nb)[0]), offs[:nb + 1]))
    data, soffs, nblk = container.pack_batch(files)
    streams = [data[int(soffs[i]):int(soffs[i + 1])] for i in range(len(files))]
    with capi.Context(44100, 2, 2048, max_blocks=int(nblk.max())) as ctx:
        pcm, status, bits = ctx.decode_streams(streams, [int(n) for n in nblk], fmt="pcm16", want_bits=True)
    for i, (ref, offs) in enumerate(refs):
        nb = int(nblk[i])
        assert status[i] == -1
        assert np.abs(pcm[i, :nb * 2048].astype(np.int32) - ref.astype(np.int32)).max() <= 1
        assert np.array_equal(container.block_index(bits[i, :nb]), offs)


@pytest.mark.gpu
@pytest.mark.parametrize("block_size,n_chan,seed", [(2048, 2, 3), (256, 1, 4), (1024, 2, 5), (8192, 2, 6), (512, 3, 7), (4096, 1, 8)])
def test_index_prepass_grammar_vectors(block_size, n_chan, seed):
    """ULCB_IndexStreams (the token-length walk alone) against the oracle decoder's bit counts on the hand-built grammar vectors: every
    block size, 1-3 channels, every token and layout; then with damaged and truncated streams (first unlocatable block = the oracle's)."""
    from oracle import oracle
    from tests import grammar
    from ulc_codec_b200 import capi
    nb = 24
    st, sz = grammar.streams(48, nb, block_size, n_chan, seed=seed)
    with capi.Context(44100, n_chan, block_size, max_blocks=nb) as ctx:
        bb, status = ctx.index_streams(st, nb)
        assert (status == -1).all()
        for i in range(len(st)):
            assert np.array_equal(bb[i], sz[i]), i
        # damage: overshooting runs, truncation
        rng = np.random.default_rng(seed)
        bad = []
        for i, s in enumerate(st):
            s = s.copy()
            if i % 3 == 0:
                s = s[:int(rng.integers(1, len(s)))]
            elif i % 3 == 1:
                a = int(rng.integers(0, len(s) - 8)); s[a:a + 8] = 0xF8          # 8h,Fh,... noise runs that overshoot their sub-block sooner or later
            else:
                s[int(rng.integers(0, len(s)))] ^= 0x5A
            bad.append(s)
        bb, status = ctx.index_streams(bad, nb)
    for i, s in enumerate(bad):
        _, bits = oracle.Decoder(n_chan, block_size).decode_stream(np.concatenate([s, np.zeros(64, np.uint8)]), nb)
        want = (bits.astype(np.int64) + 7) // 8
        ends = np.cumsum(want)
        good = int(np.flatnonzero(bits == 0)[0]) if (bits == 0).any() else nb
        good = min(good, int((ends[:good] <= len(s)).sum()))                     # blocks that lie wholly inside a truncated stream
        got_good = nb if status[i] < 0 else int(status[i])
        assert got_good == good, (i, got_good, good)
        assert np.array_equal(bb[i, :good], want[:good]), i
        assert not bb[i, good:].any(), i


@pytest.mark.gpu
def test_index_prepass_on_ulc_files():
    """container.index_batch: golden .ulc files of one shape get their block offsets without being decoded."""
    from ulc_codec_b200 import capi
    files = []
    for name in ("music_cbr128_bs2048", "transient_vbr50_bs2048", "music_vbr75_bs2048"):
        g = load_golden_stream(name)
        files.append(container.parse(container.build(2048, 44100, 2, g["data"], g["block_bytes"])))
    with capi.Context(44100, 2, 2048, max_blocks=max(f.header.n_blocks for f in files)) as ctx:
        status = container.index_batch(ctx, files)
    assert (status == -1).all()
    for f, name in zip(files, ("music_cbr128_bs2048", "transient_vbr50_bs2048", "music_vbr75_bs2048")):
        g = load_golden_stream(name)
        assert np.array_equal(np.diff(f.offsets), g["block_bytes"].astype(np.int64))
