"""The `.ulc` container around the block stream.

Mirrors what the reference tools read and write (tools/ulc_Helper.h:10-20, tools/ulcEncodeTool.c:95-100,117-125,
161-166,190-195, tools/ulcDecodeTool.c:73-89,127-128): a 24-byte little-endian header followed (at `StreamOffs`)
by the raw, byte-aligned blocks back to back.  The format carries no block index; `block_index` derives one from
the per-block bit counts the batched decoder returns (`ULCB_DecodeStreams(..., blockBits)`), and `UlcFile.offsets`
keeps it beside the payload so batches of files can be packed for `ULCB_DecodeStreams` / written after
`ULCB_EncodeStreams` without a host-side parse.

Host logic only: nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Iterable, Optional, Sequence

import numpy as np

HEADER_MAGIC = 0x32434C55                   # 'U' | 'L'<<8 | 'C'<<16 | '2'<<24  (tools/ulc_Helper.h:10)
HEADER_STRUCT = struct.Struct("<IHHIIHHI")  # Magic, BlockSize, MaxBlockSize, nBlocks, RateHz, nChan, RateKbps, StreamOffs
HEADER_BYTES = HEADER_STRUCT.size           # 24


class UlcFormatError(ValueError):
    pass


@dataclasses.dataclass
class UlcHeader:
    block_size: int
    n_blocks: int
    rate_hz: int
    n_chan: int
    max_block_size: int = 0                 # largest block in bytes; 0 = unknown (tools/ulc_Helper.h:14)
    rate_kbps: int = 0                      # nominal coding rate, lrint of the average (tools/ulcEncodeTool.c:190)
    stream_offs: int = HEADER_BYTES

    def pack(self) -> bytes:
        for name, lim in (("block_size", 0xFFFF), ("max_block_size", 0xFFFF), ("n_chan", 0xFFFF), ("rate_kbps", 0xFFFF),
                          ("n_blocks", 0xFFFFFFFF), ("rate_hz", 0xFFFFFFFF), ("stream_offs", 0xFFFFFFFF)):
            v = getattr(self, name)
            if not 0 <= v <= lim:
                raise UlcFormatError(f"{name}={v} does not fit the header field")
        return HEADER_STRUCT.pack(HEADER_MAGIC, self.block_size, self.max_block_size, self.n_blocks, self.rate_hz,
                                  self.n_chan, self.rate_kbps, self.stream_offs)

    @classmethod
    def unpack(cls, raw: bytes) -> "UlcHeader":
        if len(raw) < HEADER_BYTES:
            raise UlcFormatError("file shorter than the 24-byte header")
        magic, bs, mbs, nb, hz, ch, kbps, offs = HEADER_STRUCT.unpack_from(raw)
        if magic != HEADER_MAGIC:                                   # tools/ulcDecodeTool.c:73
            raise UlcFormatError("bad magic: not a ULC2 file")
        if bs < 256 or bs > 32768 or bs & (bs - 1) or ch < 1 or ch > 255 or hz < 1:      # libulc/ulcDecoder.c:33-36
            raise UlcFormatError(f"unsupported parameters: BlockSize {bs}, nChan {ch}, RateHz {hz}")
        if offs < HEADER_BYTES:
            raise UlcFormatError(f"StreamOffs {offs} points inside the header")
        return cls(block_size=bs, n_blocks=nb, rate_hz=hz, n_chan=ch, max_block_size=mbs, rate_kbps=kbps, stream_offs=offs)

    @property
    def n_samples_decoded(self) -> int:
        """Samples the reference decoder writes: every block, including the two lead-in blocks (tools/ulcDecodeTool.c:132-172)."""
        return self.n_blocks * self.block_size


def n_blocks_for(n_samples: int, block_size: int) -> int:
    """tools/ulcEncodeTool.c:98: two extra blocks flush the encoder's 2-block delay."""
    return (n_samples + block_size - 1) // block_size + 2


@dataclasses.dataclass
class UlcFile:
    header: UlcHeader
    payload: np.ndarray                      # uint8, the block stream (from StreamOffs to the end of the file)
    offsets: Optional[np.ndarray] = None     # int64 [n_blocks + 1] byte offsets of the blocks inside `payload`, if known

    def block(self, i: int) -> np.ndarray:
        if self.offsets is None:
            raise UlcFormatError("no block index: call block_index() with the decoder's per-block bit counts first")
        return self.payload[self.offsets[i]:self.offsets[i + 1]]


def parse(raw: bytes | np.ndarray) -> UlcFile:
    raw = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else raw
    hdr = UlcHeader.unpack(raw[:HEADER_BYTES].tobytes())
    if hdr.stream_offs > len(raw):
        raise UlcFormatError("StreamOffs beyond the end of the file")
    return UlcFile(hdr, raw[hdr.stream_offs:])


def read(path: str) -> UlcFile:
    return parse(np.fromfile(path, dtype=np.uint8))


def block_index(block_bits: Sequence[int] | np.ndarray) -> np.ndarray:
    """Byte offsets of the blocks from the bits each `ULC_DecodeBlock` consumed: blocks are byte aligned
    (libulc/ulcEncoder_Encode.c:357-358, tools/ulcDecodeTool.c:160-164).  A 0 (undecodable block) ends the index."""
    bits = np.asarray(block_bits, dtype=np.int64)
    bad = np.flatnonzero(bits <= 0)
    if len(bad):
        bits = bits[:bad[0]]
    return np.concatenate([[0], np.cumsum((bits + 7) // 8)]).astype(np.int64)


def build(block_size: int, rate_hz: int, n_chan: int, payload: np.ndarray, block_bytes: Sequence[int] | np.ndarray) -> bytes:
    """Header + payload exactly as tools/ulcEncodeTool.c writes them: MaxBlockSize = largest block, RateKbps =
    lrint(total bytes * 8 * RateHz / 1000 / (BlockSize * nBlocks)) (:165,:174,:190), StreamOffs = 24."""
    bb = np.asarray(block_bytes, dtype=np.int64)
    total = int(bb.sum())
    if total != len(payload):
        raise UlcFormatError(f"block sizes sum to {total} bytes but the payload holds {len(payload)}")
    n_blocks = len(bb)
    avg_kbps = total * 8.0 * rate_hz / 1000.0 / (block_size * n_blocks) if n_blocks else 0.0
    hdr = UlcHeader(block_size=block_size, n_blocks=n_blocks, rate_hz=rate_hz, n_chan=n_chan,
                    max_block_size=int(bb.max()) if n_blocks else 0, rate_kbps=int(np.rint(avg_kbps)))   # rint = lrint (half-even)
    return hdr.pack() + np.ascontiguousarray(payload, dtype=np.uint8).tobytes()


def write(path: str, block_size: int, rate_hz: int, n_chan: int, payload: np.ndarray, block_bytes) -> None:
    with open(path, "wb") as f:
        f.write(build(block_size, rate_hz, n_chan, payload, block_bytes))


def pack_batch(files: Iterable[UlcFile]):
    """Concatenate the payloads of same-shaped files for `ULCB_DecodeStreams`: (bytes, streamOffs u64 [n+1], nBlocks u32 [n])."""
    files = list(files)
    if not files:
        raise UlcFormatError("empty batch")
    h0 = files[0].header
    for f in files:
        h = f.header
        if (h.block_size, h.n_chan, h.rate_hz) != (h0.block_size, h0.n_chan, h0.rate_hz):
            raise UlcFormatError("a batch shares BlockSize, nChan and RateHz (one ULCB_Context)")
    offs = np.zeros(len(files) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(f.payload) for f in files])
    data = np.concatenate([f.payload for f in files]) if offs[-1] else np.zeros(0, np.uint8)
    return data, offs, np.array([f.header.n_blocks for f in files], dtype=np.uint32)


def index_batch(ctx, files: Iterable[UlcFile]) -> np.ndarray:
    """The light pre-pass for foreign files: block sizes of every file of a batch from `ULCB_IndexStreams` (the
    token-length walk alone, no decode), stored as `offsets` on each file.  Returns status int32[n]: -1 = every block located, else
    the index of the first block that was not (the reference tool stops there with "Corrupted stream", tools/ulcDecodeTool.c:153-157).
    The sizes feed `ULCB_DecodeStreamsIndexed` and random access into the stream."""
    files = list(files)
    data, offs, nblk = pack_batch(files)
    streams = [data[int(offs[i]):int(offs[i + 1])] for i in range(len(files))]
    bb, status = ctx.index_streams(streams, nblk)
    for f, row, st in zip(files, bb, status):
        n = f.header.n_blocks if st < 0 else int(st)
        f.offsets = np.concatenate([[0], np.cumsum(row[:n].astype(np.int64))]).astype(np.int64)
    return status
