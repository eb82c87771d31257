"""RIFF/WAVE framing around the codec path: what tools/WavIO_Reader.c, tools/WavIO_Writer.c and
tools/MiniRIFF.c do for the reference tools, for callers of the batched API.

Only the framing lives here (host side, like the reference's): the sample conversion itself - 8-bit unsigned, 16-bit, packed
24-bit and float32 to and from float with the rounding and clamping of tools/WavIO_Helper.c:31-87 - is the device kernel
behind `ULCB_ConvertSamplesDevice` (csrc/ulcb_pcm.cu), reached through `to_float_device` / `from_float_device` below.

Reader (tools/WavIO_Reader.c:62-104, tools/MiniRIFF.c:9-52): walks the chunks of the first RIFF(WAVE) list, keeps `fmt `
(its first 16 bytes) and `data`; chunk payloads are padded to 2 bytes; the number of sample points is
data size / nChannels / bytes per sample in that order of integer division.
Writer (tools/WavIO_Writer.c:15-61, tools/WavIO_Helper.c:115-148): RIFF header, a 16-byte `fmt ` chunk, one `data` chunk padded
to 2 bytes, sizes patched at the end.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT = 0x0001, 0x0003

# (format tag, bits) -> (capi format name, numpy dtype of the raw container)
_FORMATS = {
    (WAVE_FORMAT_PCM, 8): ("pcm8u", np.uint8),
    (WAVE_FORMAT_PCM, 16): ("pcm16", np.int16),
    (WAVE_FORMAT_PCM, 24): ("pcm24", np.uint8),          # packed 3 bytes per sample
    (WAVE_FORMAT_IEEE_FLOAT, 32): ("f32", np.float32),
}


class WavError(ValueError):
    """Mirrors the reader's error codes (include/WavIO.h:9-14)."""


@dataclass
class WavFormat:
    tag: int
    n_channels: int
    rate_hz: int
    bits: int

    @property
    def sample_bytes(self) -> int:
        return self.bits // 8

    @property
    def kind(self) -> str:
        try:
            return _FORMATS[(self.tag, self.bits)][0]
        except KeyError:
            raise WavError("WAV_EUNSUPPORTED") from None

    def pack(self) -> bytes:
        """struct WAVE_fmt_t (include/WavIO.h:20-27)."""
        align = self.n_channels * self.sample_bytes
        return struct.pack("<HHIIHH", self.tag, self.n_channels, self.rate_hz, self.rate_hz * align, align, self.bits)


@dataclass
class WavData:
    fmt: WavFormat
    raw: np.ndarray            # the bytes of the data chunk that hold whole sample points, uint8
    n_sample_points: int

    def samples(self) -> np.ndarray:
        """The data chunk viewed in its container type: [nSamplePoints, nChannels] (PCM24: [.., nChannels, 3] bytes)."""
        kind, dt = _FORMATS[(self.fmt.tag, self.fmt.bits)]
        if kind == "pcm24":
            return self.raw.reshape(self.n_sample_points, self.fmt.n_channels, 3)
        return self.raw.view(dt).reshape(self.n_sample_points, self.fmt.n_channels)


def parse(buf: bytes) -> WavData:
    """WAV_OpenR on an in-memory file."""
    if len(buf) < 12:
        raise WavError("WAV_ENOTWAV")
    fmt = None
    data = None
    typ, size = struct.unpack_from("<4sI", buf, 0)
    if typ not in (b"RIFF", b"LIST") or buf[8:12] != b"WAVE":
        raise WavError("WAV_ENOTWAV")
    end = min(len(buf), 8 + ((size + 1) & ~1))
    pos = 12
    while pos + 8 <= end:
        ck, cs = struct.unpack_from("<4sI", buf, pos)
        beg = pos + 8
        if ck == b"fmt ":
            if beg + 16 > len(buf):
                raise WavError("WAV_EIO")
            tag, nch, rate, _, _, bits = struct.unpack_from("<HHIIHH", buf, beg)
            fmt = WavFormat(tag, nch, rate, bits)
        elif ck == b"data":
            data = (beg, cs)
        pos = beg + ((cs + 1) & ~1)
    if fmt is None or data is None:
        raise WavError("WAV_EINVALID")
    if not fmt.n_channels:
        raise WavError("WAV_EINVALID")
    if (fmt.tag, fmt.bits) not in _FORMATS:
        raise WavError("WAV_EUNSUPPORTED")
    n_points = data[1] // fmt.n_channels // fmt.sample_bytes        # WavIO_Reader.c:80-104: in this order
    nbytes = n_points * fmt.n_channels * fmt.sample_bytes
    avail = max(0, min(nbytes, len(buf) - data[0]))                 # a truncated file yields what is there (fread)
    n_points = avail // (fmt.n_channels * fmt.sample_bytes)
    raw = np.frombuffer(buf, np.uint8, n_points * fmt.n_channels * fmt.sample_bytes, data[0]).copy()
    return WavData(fmt, raw, n_points)


def read(path: str) -> WavData:
    with open(path, "rb") as f:
        return parse(f.read())


def build(fmt: WavFormat, raw) -> bytes:
    """WAV_OpenW + the data + WAV_Close: the bytes the reference writer produces for this format and payload."""
    _ = fmt.kind                                                     # unsupported formats are refused up front
    payload = np.ascontiguousarray(raw).tobytes()
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt.pack() + b"data" + struct.pack("<I", len(payload)) + payload
    if len(body) & 1:
        body += b"\0"                                                # WavIO_Helper.c:131
    return b"RIFF" + struct.pack("<I", len(body)) + body


def write(path: str, fmt: WavFormat, raw) -> None:
    with open(path, "wb") as f:
        f.write(build(fmt, raw))


# ---- sample conversion on the device (tools/WavIO_Helper.c:31-87 -> csrc/ulcb_pcm.cu) ----
def to_float_device(ctx, wav: WavData, d_raw: int, d_float: int) -> None:
    """WAV_ReadAsFloat's conversion step for a whole data chunk already in device memory: d_raw -> float32 at d_float."""
    ctx.convert_samples_device(wav.fmt.kind, True, d_raw, d_float, wav.n_sample_points * wav.fmt.n_channels)


def from_float_device(ctx, fmt: WavFormat, d_float: int, d_raw: int, n_samples: int) -> None:
    """WAV_WriteFromFloat's conversion step: float32 at d_float -> the container type at d_raw (n_samples = points x channels)."""
    ctx.convert_samples_device(fmt.kind, False, d_float, d_raw, n_samples)
