"""ctypes binding of oracle/libulcoracle.so (the plain-C restatement in oracle/ulc_oracle.c).

TEST INFRASTRUCTURE ONLY - see the header of ulc_oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg; never from ulc_codec_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libulcoracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "ulc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.ulco_dct4.argtypes = [_f32p, _f32p, C.c_int]
        L.ulco_dct2.argtypes = [_f32p, _f32p, C.c_int]
        L.ulco_imdct.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.ulco_mdct_mdst.argtypes = [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int]
        L.ulco_subblocks.argtypes = [C.c_int, C.POINTER(C.c_int * 4), C.POINTER(C.c_int)]
        L.ulco_subblocks.restype = C.c_int
        L.ulco_decoder_new.argtypes = [C.c_int, C.c_int]
        L.ulco_decoder_new.restype = C.c_void_p
        L.ulco_decoder_free.argtypes = [C.c_void_p]
        L.ulco_decoder_rng.argtypes = [C.c_void_p]
        L.ulco_decoder_rng.restype = C.c_uint32
        L.ulco_parse_block.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.POINTER(C.c_int)]
        L.ulco_parse_block.restype = C.c_int
        L.ulco_decode_block.argtypes = [C.c_void_p, _f32p, C.c_void_p]
        L.ulco_decode_block.restype = C.c_int
        L.ulco_float_to_pcm16.argtypes = [_i16p, _f32p, C.c_size_t]
        L.ulco_encoder_new.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ulco_encoder_new.restype = C.c_void_p
        L.ulco_encoder_free.argtypes = [C.c_void_p]
        L.ulco_encoder_window_ctrl.argtypes = [C.c_void_p]
        L.ulco_encoder_complexity.argtypes = [C.c_void_p]
        L.ulco_encoder_complexity.restype = C.c_float
        L.ulco_encoder_coefs.argtypes = [C.c_void_p]
        L.ulco_encoder_coefs.restype = C.POINTER(C.c_float)
        L.ulco_encoder_ranks.argtypes = [C.c_void_p]
        L.ulco_encoder_ranks.restype = C.POINTER(C.c_int)
        for nm in ("noise", "keys", "mask"):
            getattr(L, "ulco_encoder_" + nm).argtypes = [C.c_void_p]
            getattr(L, "ulco_encoder_" + nm).restype = C.POINTER(C.c_float)
        L.ulco_encoder_n_nz.argtypes = [C.c_void_p]
        L.ulco_encoder_n_out.argtypes = [C.c_void_p]
        L.ulco_encode_block.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_float, C.c_float, C.POINTER(C.POINTER(C.c_uint8))]
        L.ulco_encode_block.restype = C.c_int
        _lib = L
    return _lib


def dct4(x: np.ndarray) -> np.ndarray:
    buf = np.ascontiguousarray(x, dtype=np.float32).copy()
    lib().ulco_dct4(buf, np.empty_like(buf), len(buf))
    return buf


def dct2(x: np.ndarray) -> np.ndarray:
    buf = np.ascontiguousarray(x, dtype=np.float32).copy()
    lib().ulco_dct2(buf, np.empty_like(buf), len(buf))
    return buf


def imdct(coefs: np.ndarray, lap: np.ndarray, overlap: int) -> np.ndarray:
    """Returns out[n]; `lap` (n/2) is updated in place."""
    n = len(coefs)
    out = np.empty(n, np.float32)
    lib().ulco_imdct(out, np.ascontiguousarray(coefs, np.float32), lap, np.empty(n, np.float32), n, overlap)
    return out


def mdct_mdst(fresh: np.ndarray, lap: np.ndarray, overlap: int):
    """Returns (mdct[n], mdst[n]); `lap` (n) is updated in place."""
    n = len(fresh)
    a, b = np.empty(n, np.float32), np.empty(n, np.float32)
    lib().ulco_mdct_mdst(a, b, np.ascontiguousarray(fresh, np.float32), lap, np.empty(n, np.float32), n, overlap)
    return a, b


def subblocks(window_ctrl: int):
    sh = (C.c_int * 4)()
    tr = C.c_int()
    n = lib().ulco_subblocks(window_ctrl, C.byref(sh), C.byref(tr))
    return list(sh)[:n], tr.value


def float_to_pcm16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty(x.shape, np.int16)
    lib().ulco_float_to_pcm16(out.reshape(-1), x.reshape(-1), x.size)
    return out


class Decoder:
    """Streaming decoder with the semantics of a fresh ULC_DecoderState_Init (+ RNG at 1234567)."""

    def __init__(self, n_chan: int = 2, block_size: int = 2048):
        self.n_chan, self.block_size = n_chan, block_size
        self._h = lib().ulco_decoder_new(n_chan, block_size)
        if not self._h:
            raise ValueError("bad decoder parameters")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ulco_decoder_free(self._h)
            self._h = None

    def decode_block(self, buf: np.ndarray, offset: int):
        """buf must have readable slack after the block.  Returns (float32[N, nChan], bits)."""
        out = np.empty((self.block_size, self.n_chan), np.float32)
        bits = lib().ulco_decode_block(self._h, out.reshape(-1), buf.ctypes.data + offset)
        return out, bits

    def parse_block(self, buf: np.ndarray, offset: int):
        coefs = np.empty(self.n_chan * self.block_size, np.float32)
        wc = C.c_int()
        bits = lib().ulco_parse_block(self._h, buf.ctypes.data + offset, coefs, C.byref(wc))
        return coefs, wc.value, bits

    def decode_stream(self, data: np.ndarray, n_blocks: int):
        """Decode n_blocks consecutive blocks (tools/ulcDecodeTool.c:135-165 stepping rule).
        Returns (float32[nBlocks*N, nChan], int32[nBlocks] bits; 0 marks the corrupt block)."""
        pad = np.zeros(4 * self.block_size * self.n_chan + 64, np.uint8)
        buf = np.concatenate([np.ascontiguousarray(data, np.uint8), pad])
        out = np.zeros((n_blocks * self.block_size, self.n_chan), np.float32)
        bits = np.zeros(n_blocks, np.int32)
        pos = 0
        for b in range(n_blocks):
            if pos >= len(data):
                break
            blk, nb = self.decode_block(buf, pos)
            bits[b] = nb
            if nb == 0:
                break
            out[b * self.block_size:(b + 1) * self.block_size] = blk
            pos += (nb + 7) // 8
        return out, bits


MODES = {"vbr": 0, "cbr": 1, "abr": 2}


class Encoder:
    """Streaming encoder with the semantics of a fresh ULC_EncoderState_Init + ULC_EncodeBlock_{VBR,CBR,ABR}."""

    def __init__(self, rate_hz: int = 44100, n_chan: int = 2, block_size: int = 2048):
        self.rate_hz, self.n_chan, self.block_size = rate_hz, n_chan, block_size
        self._h = lib().ulco_encoder_new(rate_hz, n_chan, block_size)
        if not self._h:
            raise ValueError("bad encoder parameters")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ulco_encoder_free(self._h)
            self._h = None

    def encode_block(self, pcm_f32: np.ndarray, mode: str, param: float, avg_complexity: float = 0.0) -> bytes:
        out = C.POINTER(C.c_uint8)()
        bits = lib().ulco_encode_block(self._h, np.ascontiguousarray(pcm_f32, np.float32).reshape(-1), MODES[mode],
                                       param, avg_complexity, C.byref(out))
        return C.string_at(out, bits // 8)

    @property
    def window_ctrl(self) -> int:
        return lib().ulco_encoder_window_ctrl(self._h)

    @property
    def complexity(self) -> float:
        return float(lib().ulco_encoder_complexity(self._h))

    def coefs(self) -> np.ndarray:
        n = self.n_chan * self.block_size
        return np.ctypeslib.as_array(lib().ulco_encoder_coefs(self._h), (n,)).copy()

    def noise(self) -> np.ndarray:
        n = self.n_chan * self.block_size
        return np.ctypeslib.as_array(lib().ulco_encoder_noise(self._h), (n,)).copy()

    def keys(self) -> np.ndarray:
        n = self.n_chan * self.block_size
        return np.ctypeslib.as_array(lib().ulco_encoder_keys(self._h), (n,)).copy()

    def mask(self) -> np.ndarray:
        return np.ctypeslib.as_array(lib().ulco_encoder_mask(self._h), (self.block_size // 2,)).copy()

    @property
    def n_nz(self) -> int:
        return lib().ulco_encoder_n_nz(self._h)

    @property
    def n_out(self) -> int:
        return lib().ulco_encoder_n_out(self._h)

    def ranks(self) -> np.ndarray:
        n = self.n_chan * self.block_size
        return np.ctypeslib.as_array(lib().ulco_encoder_ranks(self._h), (n,)).copy()

    def encode_stream(self, pcm16: np.ndarray, mode: str, param: float, avg_complexity: float = 0.0, dump: bool = False):
        """pcm16: int16[nSamples, nChan]; PCM16 -> float by x*2^-15 (tools/WavIO_Helper.c:49-55), zero padded,
        nBlocks = ceil(n/N)+2 (tools/ulcEncodeTool.c:98).  Returns dict(data, block_bytes, window_ctrl, complexity[, coefs, ranks])."""
        N, Cn = self.block_size, self.n_chan
        ns = len(pcm16)
        nb = (ns + N - 1) // N + 2
        x = np.zeros((nb * N, Cn), np.float32)
        x[:ns] = pcm16.astype(np.float32) * np.float32(2.0 ** -15)
        chunks, sizes, wcs, cxs, coefs, ranks = [], [], [], [], [], []
        extra = {}
        for b in range(nb):
            by = self.encode_block(x[b * N:(b + 1) * N], mode, param, avg_complexity)
            chunks.append(by); sizes.append(len(by)); wcs.append(self.window_ctrl); cxs.append(self.complexity)
            if dump:
                coefs.append(self.coefs()); ranks.append(self.ranks())
                if dump == "all":
                    extra.setdefault("noise", []).append(self.noise()); extra.setdefault("keys", []).append(self.keys())
                    extra.setdefault("mask", []).append(self.mask()); extra.setdefault("n_nz", []).append(self.n_nz)
                    extra.setdefault("n_out", []).append(self.n_out)
        r = dict(data=np.frombuffer(b"".join(chunks), np.uint8).copy(), block_bytes=np.array(sizes, np.uint32),
                 window_ctrl=np.array(wcs, np.int32), complexity=np.array(cxs, np.float32))
        if dump:
            r["coefs"] = np.stack(coefs); r["ranks"] = np.stack(ranks)
            for k_, v_ in extra.items():
                r[k_] = np.stack(v_) if np.ndim(v_[0]) else np.array(v_)
        return r


# ---- sample formats (tools/WavIO_Helper.c:31-87), numpy restatement; pinned against the reference tools in
# ---- tests/test_sample_formats.py (ulcdecodetool -format:..., ulcencodetool reading 8/16/24-bit and float WAVs)
def samples_to_float(raw: np.ndarray, fmt: str) -> np.ndarray:
    """raw: uint8 (pcm8u), int16 (pcm16), uint8[3n] little-endian packed (pcm24) or float32 (f32) -> float32."""
    if fmt == "pcm8u":
        # :31-37 as written: the byte is read as int8, PROMOTED to int and only then xor-ed with 0x80, so the result is
        # b+128 for b < 128 and b-384 above (outside +-1.0) - reproduced, not repaired, because that is what the
        # reference encoder sees for an 8-bit WAV
        return (raw.astype(np.uint8).view(np.int8).astype(np.int32) ^ 0x80).astype(np.float32) * np.float32(2.0 ** -7)
    if fmt == "pcm16":
        return raw.astype(np.int16).astype(np.float32) * np.float32(2.0 ** -15)                                # :49-55
    if fmt == "pcm24":
        b = raw.astype(np.uint8).reshape(-1, 3).astype(np.uint32)
        x = ((b[:, 0] << 8) | (b[:, 1] << 16) | (b[:, 2] << 24)).astype(np.uint32).view(np.int32)
        return x.astype(np.float32) * np.float32(2.0 ** -31)                                                    # :67-77
    if fmt == "f32":
        return raw.astype(np.float32)
    raise ValueError(fmt)


def samples_from_float(x: np.ndarray, fmt: str) -> np.ndarray:
    """float32 -> the format's storage (uint8 / int16 / packed uint8[3n] / float32): clamp, lrintf (half-even)."""
    x = x.astype(np.float32)
    if fmt == "pcm8u":
        q = np.rint(np.clip(x * np.float32(128.0), np.float32(-128.0), np.float32(127.0))).astype(np.int8)    # :39-45
        return q.view(np.uint8) ^ np.uint8(0x80)
    if fmt == "pcm16":
        return float_to_pcm16(x)
    if fmt == "pcm24":
        q = np.rint(np.clip(x * np.float32(8388608.0), np.float32(-8388608.0), np.float32(8388607.0))).astype(np.int32).view(np.uint32)
        out = np.empty((len(q), 3), np.uint8)
        out[:, 0] = q & 0xFF; out[:, 1] = (q >> 8) & 0xFF; out[:, 2] = (q >> 16) & 0xFF                         # :79-87
        return out.reshape(-1)
    if fmt == "f32":
        return x
    raise ValueError(fmt)
