"""ctypes binding of libulcb200.so (include/ulcb200.h) - the host-side mirror of the C ABI.

The product path is the CUDA library; this module only marshals numpy / torch buffers into the
plain-pointer C calls.  It fails loudly (ImportError / RuntimeError) when the library is missing or
no sm_100 GPU is present - there is no CPU fallback and nothing here touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libulcb200.so")

FMT_PCM16, FMT_FLOAT32 = 0, 1
FMT_PCM8U, FMT_PCM24 = 2, 3
_FMT = {"pcm16": FMT_PCM16, "s16": FMT_PCM16, "f32": FMT_FLOAT32, "float32": FMT_FLOAT32, "pcm8u": FMT_PCM8U, "pcm24": FMT_PCM24}

# every symbol include/ulcb200.h declares (tests/test_capi_symbols.py checks the header against this)
EXPORTS = (
    "ULC_DecoderState_Init", "ULC_DecoderState_Destroy", "ULC_DecodeBlock",
    "ULCB_ContextCreate", "ULCB_ContextDestroy", "ULCB_LastCudaError", "ULCB_ErrorString",
    "ULCB_MaxBlockBytes", "ULCB_DecodeStreamsDevice", "ULCB_DecodeStreams", "ULCB_DecodeStreamsIndexedDevice",
    "ULCB_DecodeStreamsIndexed", "ULCB_IndexStreams", "ULCB_Sync", "ULCB_Stream",
    "ULCB_LaunchCount", "ULCB_SetKernelTiming", "ULCB_KernelTimes", "ULCB_SetDecodePipeline", "ULCB_GetDecodePipeline", "ULCB_TestDCT4", "ULCB_TestParse",
    "ULC_EncoderState_Init", "ULC_EncoderState_Destroy", "ULC_EncodeBlock_CBR", "ULC_EncodeBlock_ABR", "ULC_EncodeBlock_VBR",
    "ULCB_BlocksForSamples", "ULCB_EncodeStreamsDevice", "ULCB_EncodeStreams", "ULCB_TestEncode",
    "ULCB_GenerateCorpusDevice", "ULCB_ConvertSamplesDevice", "ULCB_GetAvgComplexity",
)

HAS_BLOCK_INDEX = True          # ULCB_DecodeStreamsIndexed[Device]: decode with a caller-supplied block-size index

MODE_VBR, MODE_CBR, MODE_ABR = 0, 1, 2
MODE_ABR_2PASS = 3
_MODE = {"vbr": MODE_VBR, "cbr": MODE_CBR, "abr": MODE_ABR, "abr2": MODE_ABR_2PASS}


class ULCB_EncodeParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("rate", C.c_float), ("avgComplexity", C.c_float)]


class ULC_EncoderState_t(C.Structure):
    """include/ulcb200.h (layout of /root/reference/include/ulcEncoder.h:47-78)."""
    _fields_ = [("RateHz", C.c_int), ("nChan", C.c_int), ("BlockSize", C.c_int),
                ("WindowCtrl", C.c_int), ("NextWindowCtrl", C.c_int), ("BlockComplexity", C.c_float),
                ("TransientFilter", C.c_float * 3), ("BufferData", C.c_void_p),
                ("SampleBuffer", C.c_void_p), ("TransformBuffer", C.c_void_p), ("TransformNoise", C.c_void_p),
                ("TransformFwdLap", C.c_void_p), ("TransformTemp", C.c_void_p), ("TransformIndex", C.c_void_p),
                ("TransientBuffer", C.c_void_p)]


class ULC_DecoderState_t(C.Structure):
    """include/ulcb200.h (layout of /root/reference/include/ulcDecoder.h:13-32)."""
    _fields_ = [("nChan", C.c_int), ("BlockSize", C.c_int), ("LastSubBlockSize", C.c_int),
                ("BufferData", C.c_void_p), ("TransformBuffer", C.c_void_p),
                ("TransformTemp", C.c_void_p), ("TransformInvLap", C.c_void_p)]


_lib = None


def load():
    """dlopen the library and declare prototypes.  Needs no GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing - run `python __graft_entry__.py` (nvcc, sm_100a) first; "
                          "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        getattr(L, name)  # AttributeError if a declared symbol is not exported
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.ULCB_ContextCreate.argtypes = [C.POINTER(vp), i32, i32, i32, i32, u32]
    L.ULCB_ContextDestroy.argtypes = [vp]
    L.ULCB_ContextDestroy.restype = None
    L.ULCB_LastCudaError.argtypes = [vp]
    L.ULCB_ErrorString.argtypes = [i32]
    L.ULCB_ErrorString.restype = C.c_char_p
    L.ULCB_MaxBlockBytes.argtypes = [i32, i32]
    L.ULCB_MaxBlockBytes.restype = C.c_size_t
    L.ULCB_DecodeStreamsDevice.argtypes = [vp, i32, vp, vp, vp, u32, vp, u64, i32, vp, vp, u64]
    L.ULCB_DecodeStreams.argtypes = [vp, i32, vp, vp, vp, u32, vp, u64, i32, vp, vp, u64]
    L.ULCB_DecodeStreamsIndexedDevice.argtypes = [vp, i32, vp, vp, vp, u32, vp, u64, vp, u64, i32, vp, vp, u64]
    L.ULCB_DecodeStreamsIndexed.argtypes = [vp, i32, vp, vp, vp, u32, vp, u64, vp, u64, i32, vp, vp, u64]
    L.ULCB_IndexStreams.argtypes = [vp, i32, vp, vp, vp, u32, vp, u64, vp]
    L.ULCB_Sync.argtypes = [vp]
    L.ULCB_Stream.argtypes = [vp]
    L.ULCB_Stream.restype = vp
    L.ULCB_LaunchCount.argtypes = [vp]
    L.ULCB_LaunchCount.restype = u64
    L.ULCB_SetKernelTiming.argtypes = [vp, i32]
    L.ULCB_KernelTimes.argtypes = [vp, vp, vp, i32]
    L.ULCB_SetDecodePipeline.argtypes = [vp, i32, i32]
    L.ULCB_GetDecodePipeline.argtypes = [vp, vp, vp]
    L.ULCB_TestDCT4.argtypes = [vp, vp, i32, i32]
    L.ULCB_TestParse.argtypes = [vp, vp, u64, u32, vp, vp]
    L.ULC_DecoderState_Init.argtypes = [C.POINTER(ULC_DecoderState_t)]
    L.ULC_DecoderState_Destroy.argtypes = [C.POINTER(ULC_DecoderState_t)]
    L.ULC_DecoderState_Destroy.restype = None
    L.ULC_DecodeBlock.argtypes = [C.POINTER(ULC_DecoderState_t), vp, vp]
    f32, pp = C.c_float, C.POINTER(ULCB_EncodeParams)
    L.ULCB_BlocksForSamples.argtypes = [u64, i32]
    L.ULCB_BlocksForSamples.restype = u32
    L.ULCB_EncodeStreamsDevice.argtypes = [vp, i32, vp, i32, vp, u64, vp, u32, pp, vp, u64, vp, vp, vp, u64, vp, vp]
    L.ULCB_EncodeStreams.argtypes = [vp, i32, vp, i32, vp, u64, u32, pp, vp, u64, vp, vp, vp, vp, vp]
    L.ULCB_TestEncode.argtypes = [vp, i32, vp, i32, vp, u64, u32, pp, vp, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ULCB_GenerateCorpusDevice.argtypes = [vp, i32, u32, i32, u64, u64, vp, vp, vp, u64]
    L.ULCB_ConvertSamplesDevice.argtypes = [vp, i32, i32, vp, vp, u64]
    L.ULCB_GetAvgComplexity.argtypes = [vp, vp, i32]
    es = C.POINTER(ULC_EncoderState_t)
    L.ULC_EncoderState_Init.argtypes = [es]
    L.ULC_EncoderState_Destroy.argtypes = [es]
    L.ULC_EncoderState_Destroy.restype = None
    L.ULC_EncodeBlock_CBR.argtypes = [es, vp, C.POINTER(C.c_int), f32]
    L.ULC_EncodeBlock_ABR.argtypes = [es, vp, C.POINTER(C.c_int), f32, f32]
    L.ULC_EncodeBlock_VBR.argtypes = [es, vp, C.POINTER(C.c_int), f32]
    for n in ("ULC_EncodeBlock_CBR", "ULC_EncodeBlock_ABR", "ULC_EncodeBlock_VBR"):
        getattr(L, n).restype = vp
    _lib = L
    return L


class UlcbError(RuntimeError):
    pass


def _check(L, ctx, rc: int, what: str):
    if rc != 0:
        cuda = L.ULCB_LastCudaError(ctx) if ctx else 0
        raise UlcbError(f"{what}: {L.ULCB_ErrorString(rc).decode()} (status {rc}, cuda error {cuda})")


def pack_streams(streams):
    """list of uint8 arrays -> (concatenated uint8, uint64 offsets[n+1])."""
    offs = np.zeros(len(streams) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in streams], dtype=np.uint64)
    data = np.concatenate([np.ascontiguousarray(s, np.uint8) for s in streams]) if streams else np.zeros(0, np.uint8)
    return data, offs


class Context:
    """ULCB_Context wrapper.  One context = one (rate, channels, block size) configuration on one GPU."""

    def __init__(self, rate_hz: int = 44100, n_chan: int = 2, block_size: int = 2048, *,
                 max_blocks: int = 4096, device: int = 0):
        self.L = load()
        self.rate_hz, self.n_chan, self.block_size, self.max_blocks = rate_hz, n_chan, block_size, max_blocks
        h = C.c_void_p()
        rc = self.L.ULCB_ContextCreate(C.byref(h), device, rate_hz, n_chan, block_size, max_blocks)
        _check(self.L, None, rc, "ULCB_ContextCreate")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.ULCB_ContextDestroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launches(self) -> int:
        return int(self.L.ULCB_LaunchCount(self.h))

    @property
    def cuda_stream(self) -> int:
        return int(self.L.ULCB_Stream(self.h) or 0)

    def sync(self):
        _check(self.L, self.h, self.L.ULCB_Sync(self.h), "ULCB_Sync")

    KERNEL_KINDS = ("bp_walk", "bp_index", "bp_parse", "bp_scan", "bp_transform", "enc_detect", "enc_encode")

    def set_kernel_timing(self, enable: bool = True):
        _check(self.L, self.h, self.L.ULCB_SetKernelTiming(self.h, int(enable)), "ULCB_SetKernelTiming")

    def set_decode_pipeline(self, chunks: int, overlap: int):
        """Chunked pipeline of the block-parallel decode (ULCB_SetDecodePipeline); the output does not depend on it."""
        _check(self.L, self.h, self.L.ULCB_SetDecodePipeline(self.h, int(chunks), int(overlap)), "ULCB_SetDecodePipeline")

    def decode_pipeline(self) -> tuple:
        """(chunks, overlap) in effect (ULCB_GetDecodePipeline)."""
        ch, ov = C.c_int(0), C.c_int(0)
        _check(self.L, self.h, self.L.ULCB_GetDecodePipeline(self.h, C.byref(ch), C.byref(ov)), "ULCB_GetDecodePipeline")
        return ch.value, ov.value

    def kernel_times(self, reset: bool = True) -> dict:
        """{kind: (milliseconds, launches)} accumulated since the last reset (ULCB_KernelTimes; synchronises the stream)."""
        ms = np.zeros(len(self.KERNEL_KINDS), np.float64)
        cnt = np.zeros(len(self.KERNEL_KINDS), np.uint64)
        _check(self.L, self.h, self.L.ULCB_KernelTimes(self.h, ms.ctypes.data, cnt.ctypes.data, int(reset)), "ULCB_KernelTimes")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNEL_KINDS)}

    # ---- host-pointer batched decode (ULCB_DecodeStreams) ----
    def decode_streams(self, streams, n_blocks, fmt: str = "pcm16", want_bits: bool = False, block_bytes=None, out=None):
        """streams: list of uint8 arrays (raw block bytes).  n_blocks: int or per-stream sequence.
        block_bytes: optional block-size index, one uint32 sequence per stream (ULCB_DecodeStreamsIndexed).
        out: optional caller-owned output array of the returned pcm's size and dtype.
        Returns (pcm[nStreams, maxBlocks*N, nChan], status int32[nStreams], bits uint32[nStreams, maxBlocks] | None)."""
        data, offs = pack_streams(streams)
        ns = len(streams)
        if np.isscalar(n_blocks):
            nb_arr, nb_all, max_blk = None, int(n_blocks), int(n_blocks)
        else:
            nb_arr = np.ascontiguousarray(n_blocks, np.uint32)
            nb_all, max_blk = 0, int(nb_arr.max()) if ns else 0
        stride = max_blk * self.block_size * self.n_chan
        f = _FMT[fmt]
        shape, dt = (ns, max_blk * self.block_size, self.n_chan), np.int16 if f == FMT_PCM16 else np.float32
        if out is None:
            out = np.zeros(shape, dt)
        else:
            # caller's buffer, e.g. page-locked memory (the kernel then stores the PCM straight into it and the call is pipelined)
            assert out.dtype == dt and out.size == int(np.prod(shape)) and out.flags.c_contiguous
            out = out.reshape(shape)
            out[...] = 0
        status = np.zeros(ns, np.int32)
        bits = np.zeros((ns, max_blk), np.uint32) if want_bits else None
        common = (out.ctypes.data, stride, f, status.ctypes.data, bits.ctypes.data if want_bits else None, max_blk)
        if block_bytes is None:
            rc = self.L.ULCB_DecodeStreams(
                self.h, ns, data.ctypes.data, offs.ctypes.data, nb_arr.ctypes.data if nb_arr is not None else None, nb_all, *common)
            _check(self.L, self.h, rc, "ULCB_DecodeStreams")
        else:
            idx = np.zeros((ns, max(1, max_blk)), np.uint32)
            for i, b in enumerate(block_bytes):
                b = np.asarray(b, np.uint32)[:max_blk]
                idx[i, :len(b)] = b
            rc = self.L.ULCB_DecodeStreamsIndexed(
                self.h, ns, data.ctypes.data, offs.ctypes.data, nb_arr.ctypes.data if nb_arr is not None else None, nb_all,
                idx.ctypes.data, idx.shape[1], *common)
            _check(self.L, self.h, rc, "ULCB_DecodeStreamsIndexed")
        return out, status, bits

    def index_streams(self, streams, n_blocks):
        """Block sizes of foreign streams (ULCB_IndexStreams: token-length walk only).  Returns (block_bytes uint32[nStreams, maxBlocks],
        status int32[nStreams]); a row is zero from the first block that could not be located."""
        data, offs = pack_streams(streams)
        ns = len(streams)
        if np.isscalar(n_blocks):
            nb_arr, nb_all, max_blk = None, int(n_blocks), int(n_blocks)
        else:
            nb_arr = np.ascontiguousarray(n_blocks, np.uint32)
            nb_all, max_blk = 0, int(nb_arr.max()) if ns else 0
        bb = np.zeros((ns, max(1, max_blk)), np.uint32)
        status = np.zeros(ns, np.int32)
        rc = self.L.ULCB_IndexStreams(self.h, ns, data.ctypes.data, offs.ctypes.data, nb_arr.ctypes.data if nb_arr is not None else None, nb_all,
                                      bb.ctypes.data, bb.shape[1], status.ctypes.data)
        _check(self.L, self.h, rc, "ULCB_IndexStreams")
        return bb, status

    # ---- device-pointer batched decode (ULCB_DecodeStreamsDevice); pointers are plain ints ----
    def decode_streams_device(self, n_streams: int, d_bytes: int, d_offs: int, n_blocks_all: int, d_out: int,
                              out_stride: int, fmt: str = "pcm16", d_status: int = 0, d_nblocks: int = 0,
                              d_bits: int = 0, bits_stride: int = 0, d_block_bytes: int = 0, block_bytes_stride: int = 0):
        if d_block_bytes:
            rc = self.L.ULCB_DecodeStreamsIndexedDevice(self.h, n_streams, d_bytes, d_offs, d_nblocks or None, n_blocks_all,
                                                        d_block_bytes, block_bytes_stride, d_out, out_stride, _FMT[fmt],
                                                        d_status or None, d_bits or None, bits_stride)
            _check(self.L, self.h, rc, "ULCB_DecodeStreamsIndexedDevice")
            return
        rc = self.L.ULCB_DecodeStreamsDevice(self.h, n_streams, d_bytes, d_offs, d_nblocks or None, n_blocks_all,
                                             d_out, out_stride, _FMT[fmt], d_status or None, d_bits or None, bits_stride)
        _check(self.L, self.h, rc, "ULCB_DecodeStreamsDevice")

    # ---- host-pointer batched encode (ULCB_EncodeStreams) ----
    def blocks_for(self, n_samples: int) -> int:
        return int(self.L.ULCB_BlocksForSamples(int(n_samples), self.block_size))

    def encode_streams(self, pcm_list, mode: str, rate: float, avg_complexity: float = 0.0, *, n_blocks: int | None = None,
                       taps: bool = False, out_stride: int | None = None):
        """pcm_list: list of int16 (or float32) arrays [nSamples, nChan] (all the same dtype).
        Every stream is coded to n_blocks blocks (default: the tool's ceil(n/N)+2 of the longest stream), zero padded.
        Returns dict(streams=[uint8 payload per stream], block_bytes[ns, nb], complexity[ns, nb], window_ctrl[ns, nb],
        status[ns]) plus, with taps=True, coefs/keys/noise [ns, nb, nChan*N], mask [ns, nb, N/2], ranks, counts [ns, nb, 2]."""
        ns, N, Cn = len(pcm_list), self.block_size, self.n_chan
        isf = np.asarray(pcm_list[0]).dtype == np.float32
        dt = np.float32 if isf else np.int16
        lens = np.array([len(p) for p in pcm_list], np.uint64)
        nb = int(n_blocks) if n_blocks else self.blocks_for(int(lens.max()))
        stride = max(1, int(lens.max())) * Cn
        buf = np.zeros((ns, stride), dt)
        for i, p in enumerate(pcm_list):
            a = np.ascontiguousarray(p, dt).reshape(-1)
            buf[i, :len(a)] = a
        ostride = int(out_stride) if out_stride else nb * int(self.L.ULCB_MaxBlockBytes(Cn, N))
        out = np.zeros((ns, ostride), np.uint8)
        bb = np.zeros((ns, nb), np.uint32); cx = np.zeros((ns, nb), np.float32); wc = np.zeros((ns, nb), np.int32)
        tot = np.zeros(ns, np.uint64); status = np.zeros(ns, np.int32)
        prm = ULCB_EncodeParams(_MODE[mode], float(rate), float(avg_complexity))
        r = {}
        if taps:
            per = Cn * N
            t = dict(coefs=np.zeros((ns, nb, per), np.float32), keys=np.zeros((ns, nb, per), np.float32),
                     noise=np.zeros((ns, nb, per), np.float32), mask=np.zeros((ns, nb, N // 2), np.float32),
                     ranks=np.zeros((ns, nb, per), np.int32), counts=np.zeros((ns, nb, 2), np.int32))
            rc = self.L.ULCB_TestEncode(self.h, ns, buf.ctypes.data, FMT_FLOAT32 if isf else FMT_PCM16, lens.ctypes.data, stride, nb,
                                        C.byref(prm), out.ctypes.data, ostride, bb.ctypes.data, cx.ctypes.data, wc.ctypes.data,
                                        *(t[k].ctypes.data for k in ("coefs", "keys", "noise", "mask", "ranks", "counts")))
            _check(self.L, self.h, rc, "ULCB_TestEncode")
            tot = bb.sum(axis=1).astype(np.uint64)
            r.update(t)
        else:
            rc = self.L.ULCB_EncodeStreams(self.h, ns, buf.ctypes.data, FMT_FLOAT32 if isf else FMT_PCM16, lens.ctypes.data, stride, nb,
                                           C.byref(prm), out.ctypes.data, ostride, bb.ctypes.data, cx.ctypes.data, wc.ctypes.data,
                                           tot.ctypes.data, status.ctypes.data)
            _check(self.L, self.h, rc, "ULCB_EncodeStreams")
        r.update(streams=[out[i, :int(tot[i])].copy() for i in range(ns)], block_bytes=bb, complexity=cx, window_ctrl=wc, status=status)
        return r

    def encode_streams_device(self, n_streams: int, d_pcm: int, fmt: str, in_stride: int, n_blocks_all: int, mode: str, rate: float,
                              d_out: int, out_stride: int, *, avg_complexity: float = 0.0, d_nsamples: int = 0, d_nblocks: int = 0,
                              d_block_bytes: int = 0, d_complexity: int = 0, d_window_ctrl: int = 0, meta_stride: int = 0,
                              d_total: int = 0, d_status: int = 0):
        prm = ULCB_EncodeParams(_MODE[mode], float(rate), float(avg_complexity))
        rc = self.L.ULCB_EncodeStreamsDevice(self.h, n_streams, d_pcm, _FMT[fmt], d_nsamples or None, in_stride, d_nblocks or None,
                                             n_blocks_all, C.byref(prm), d_out, out_stride, d_block_bytes or None, d_complexity or None,
                                             d_window_ctrl or None, meta_stride, d_total or None, d_status or None)
        _check(self.L, self.h, rc, "ULCB_EncodeStreamsDevice")

    def generate_corpus_device(self, kind: str, first_id: int, n_streams: int, n_samples: int, d_out: int, stride_elems: int, t0: int = 0):
        """Fill device memory with the synthetic corpus (bit-identical to corpus.generate)."""
        from . import corpus
        sine, burst = corpus.sine_lut(), corpus.burst_lut()
        rc = self.L.ULCB_GenerateCorpusDevice(self.h, corpus._KINDS[kind], first_id, n_streams, n_samples, t0,
                                              sine.ctypes.data, burst.ctypes.data, d_out, stride_elems)
        _check(self.L, self.h, rc, "ULCB_GenerateCorpusDevice")

    def avg_complexity(self, n_streams: int) -> np.ndarray:
        """Per-stream AvgComplexity measured by the last mode="abr2" (two-pass ABR) call."""
        out = np.zeros(n_streams, np.float32)
        _check(self.L, self.h, self.L.ULCB_GetAvgComplexity(self.h, out.ctypes.data, n_streams), "ULCB_GetAvgComplexity")
        return out

    def convert_samples_device(self, fmt: str, to_float: bool, d_src: int, d_dst: int, n_samples: int):
        """Device-side sample-format conversion with the reference's rules (tools/WavIO_Helper.c:31-87).
        fmt in {"pcm8u", "pcm16", "pcm24", "f32"}; the other side is float32."""
        rc = self.L.ULCB_ConvertSamplesDevice(self.h, _FMT[fmt], 1 if to_float else 0, d_src, d_dst, n_samples)
        _check(self.L, self.h, rc, "ULCB_ConvertSamplesDevice")

    def test_parse(self, data: np.ndarray, n_blocks: int):
        """(coefs float32[nBlocks, nChan*N], bits uint32[nBlocks]) of one stream: parse + dequantise + noise only."""
        data = np.ascontiguousarray(data, np.uint8)
        coefs = np.zeros((n_blocks, self.n_chan * self.block_size), np.float32)
        bits = np.zeros(n_blocks, np.uint32)
        rc = self.L.ULCB_TestParse(self.h, data.ctypes.data, len(data), n_blocks, coefs.ctypes.data, bits.ctypes.data)
        _check(self.L, self.h, rc, "ULCB_TestParse")
        return coefs, bits

    def test_dct4(self, x: np.ndarray) -> np.ndarray:
        buf = np.ascontiguousarray(x, np.float32).copy()
        count, size = (1, buf.shape[0]) if buf.ndim == 1 else buf.shape
        _check(self.L, self.h, self.L.ULCB_TestDCT4(self.h, buf.ctypes.data, size, count), "ULCB_TestDCT4")
        return buf


class DropinDecoder:
    """The reference's single-stream API (ULC_DecoderState_Init / ULC_DecodeBlock), through the drop-in symbols."""

    def __init__(self, n_chan: int = 2, block_size: int = 2048):
        self.L = load()
        self.st = ULC_DecoderState_t()
        self.st.nChan, self.st.BlockSize = n_chan, block_size
        self.ok = self.L.ULC_DecoderState_Init(C.byref(self.st))
        if self.ok < 0:
            raise UlcbError("ULC_DecoderState_Init returned -1")

    def decode_block(self, buf: np.ndarray, offset: int = 0):
        out = np.empty((self.st.BlockSize, self.st.nChan), np.float32)
        bits = self.L.ULC_DecodeBlock(C.byref(self.st), out.ctypes.data, buf.ctypes.data + offset)
        return out, bits

    def close(self):
        if getattr(self, "ok", -1) >= 0:
            self.L.ULC_DecoderState_Destroy(C.byref(self.st))
            self.ok = -1

    __del__ = close


class DropinEncoder:
    """The reference's single-stream API (ULC_EncoderState_Init / ULC_EncodeBlock_*), through the drop-in symbols."""

    def __init__(self, rate_hz: int = 44100, n_chan: int = 2, block_size: int = 2048):
        self.L = load()
        self.st = ULC_EncoderState_t()
        self.st.RateHz, self.st.nChan, self.st.BlockSize = rate_hz, n_chan, block_size
        self.ok = self.L.ULC_EncoderState_Init(C.byref(self.st))
        if self.ok < 0:
            raise UlcbError("ULC_EncoderState_Init returned -1")

    def encode_block(self, pcm_f32: np.ndarray, mode: str, rate: float, avg_complexity: float = 0.0) -> bytes:
        x = np.ascontiguousarray(pcm_f32, np.float32)
        size = C.c_int(0)
        if mode == "vbr":
            p = self.L.ULC_EncodeBlock_VBR(C.byref(self.st), x.ctypes.data, C.byref(size), rate)
        elif mode == "abr":
            p = self.L.ULC_EncodeBlock_ABR(C.byref(self.st), x.ctypes.data, C.byref(size), rate, avg_complexity)
        else:
            p = self.L.ULC_EncodeBlock_CBR(C.byref(self.st), x.ctypes.data, C.byref(size), rate)
        if not p:
            raise UlcbError("ULC_EncodeBlock_* returned NULL")
        return C.string_at(p, (size.value + 7) // 8)

    def close(self):
        if getattr(self, "ok", -1) >= 0:
            self.L.ULC_EncoderState_Destroy(C.byref(self.st))
            self.ok = -1

    __del__ = close
