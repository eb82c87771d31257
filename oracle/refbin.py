"""Python driver for oracle/_ref/ulcref_harness_* (the UNMODIFIED reference, compiled by
oracle/Makefile).  TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never from ulc_codec_b200/.

One subprocess per stream: the reference decoder's RNG is a process-global static
(libulc/ulcDecoder.c:75-81), so "fresh process" is the only way to restart it at 1234567.
"""
from __future__ import annotations

import os
import subprocess
import tempfile
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def harness_path(variant: str = "scalar") -> str:
    return os.path.join(REF_DIR, f"ulcref_harness_{variant}")


def available(variant: str = "scalar") -> bool:
    return os.access(harness_path(variant), os.X_OK)


@dataclass
class RefEncoded:
    data: np.ndarray          # uint8, concatenated byte-aligned blocks
    block_bytes: np.ndarray   # uint32[nBlocks]
    window_ctrl: np.ndarray   # int32[nBlocks]
    complexity: np.ndarray    # float32[nBlocks]
    coefs: np.ndarray | None  # float32[nBlocks, nChan*N]  (State.TransformBuffer)
    ranks: np.ndarray | None  # int32[nBlocks, nChan*N]    (State.TransformIndex)


def encode(pcm16: np.ndarray, mode: str, param: float, *, avg_complexity: float = 0.0,
           block_size: int = 2048, rate_hz: int = 44100, variant: str = "scalar",
           dump: bool = False) -> RefEncoded:
    """pcm16: int16[nSamples, nChan].  mode in {"vbr","cbr","abr"}; param = Quality or kbps."""
    pcm16 = np.ascontiguousarray(pcm16, dtype=np.int16)
    n_chan = pcm16.shape[1]
    with tempfile.TemporaryDirectory() as td:
        p = lambda n: os.path.join(td, n)
        pcm16.tofile(p("in.s16"))
        cmd = [harness_path(variant), "enc", mode, repr(float(param)), repr(float(avg_complexity)),
               str(block_size), str(n_chan), str(rate_hz), p("in.s16"), p("o.bytes"), p("o.meta")]
        if dump:
            cmd += [p("o.coefs"), p("o.ranks")]
        subprocess.run(cmd, check=True)
        data = np.fromfile(p("o.bytes"), dtype=np.uint8)
        meta = np.fromfile(p("o.meta"), dtype=np.dtype([("nb", "<u4"), ("wc", "<i4"), ("cx", "<f4"), ("z", "<u4")]))
        coefs = ranks = None
        if dump:
            coefs = np.fromfile(p("o.coefs"), dtype=np.float32).reshape(len(meta), -1)
            ranks = np.fromfile(p("o.ranks"), dtype=np.int32).reshape(len(meta), -1)
    return RefEncoded(data, meta["nb"].copy(), meta["wc"].copy(), meta["cx"].copy(), coefs, ranks)


def decode(data: np.ndarray, n_blocks: int, *, n_chan: int = 2, block_size: int = 2048,
           variant: str = "scalar"):
    """Returns (float32[nBlocks*N, nChan] PCM, int32[nBlocks] bits consumed per block)."""
    with tempfile.TemporaryDirectory() as td:
        p = lambda n: os.path.join(td, n)
        np.ascontiguousarray(data, dtype=np.uint8).tofile(p("in.bytes"))
        subprocess.run([harness_path(variant), "dec", str(block_size), str(n_chan), str(n_blocks),
                        p("in.bytes"), p("o.f32"), p("o.bits")], check=True)
        pcm = np.fromfile(p("o.f32"), dtype=np.float32).reshape(-1, n_chan)
        bits = np.fromfile(p("o.bits"), dtype=np.int32)
    return pcm, bits


def float_to_pcm16(x: np.ndarray) -> np.ndarray:
    """tools/WavIO_Helper.c:57-63: (int16)lrintf(clamp(x*32768, -32768, 32767)), round-half-even."""
    y = np.clip(x.astype(np.float32) * np.float32(32768.0), np.float32(-32768.0), np.float32(32767.0))
    return np.rint(y).astype(np.int16)


# ---- stage-level access to the compiled reference (libfourier entry points are public) ----
_stage = {}


def stage_lib(variant: str = "scalar"):
    """ctypes handle on oracle/_ref/libulcref_<variant>.so (Fourier_DCT4/DCT2/IMDCT/MDCT_MDST)."""
    import ctypes as C
    if variant not in _stage:
        L = C.CDLL(os.path.join(REF_DIR, f"libulcref_{variant}.so"))
        f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
        L.Fourier_DCT4.argtypes = [f32p, f32p, C.c_int]
        L.Fourier_DCT2.argtypes = [f32p, f32p, C.c_int]
        L.Fourier_IMDCT.argtypes = [f32p, f32p, f32p, f32p, C.c_int, C.c_int]
        L.Fourier_MDCT_MDST.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int]
        _stage[variant] = L
    return _stage[variant]


def _aligned(n: int, dtype=np.float32) -> np.ndarray:
    """64-byte aligned scratch (the AVX build uses aligned loads)."""
    raw = np.empty(n * np.dtype(dtype).itemsize + 64, np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n * np.dtype(dtype).itemsize].view(dtype)


def ref_dct4(x, variant="scalar"):
    buf = _aligned(len(x)); buf[:] = x
    stage_lib(variant).Fourier_DCT4(buf, _aligned(len(x)), len(x))
    return buf.copy()


def ref_dct2(x, variant="scalar"):
    buf = _aligned(len(x)); buf[:] = x
    stage_lib(variant).Fourier_DCT2(buf, _aligned(len(x)), len(x))
    return buf.copy()


def ref_imdct(coefs, lap, overlap, variant="scalar"):
    n = len(coefs)
    out, cin, l, t = _aligned(n), _aligned(n), _aligned(n // 2), _aligned(n)
    cin[:] = coefs; l[:] = lap
    stage_lib(variant).Fourier_IMDCT(out, cin, l, t, n, overlap)
    lap[:] = l
    return out.copy()


def ref_mdct_mdst(fresh, lap, overlap, variant="scalar"):
    n = len(fresh)
    a, b, f, l, t = _aligned(n), _aligned(n), _aligned(n), _aligned(n), _aligned(n)
    f[:] = fresh; l[:] = lap
    stage_lib(variant).Fourier_MDCT_MDST(a, b, f, l, t, n, overlap)
    lap[:] = l
    return a.copy(), b.copy()
