#!/usr/bin/env python
"""bench.py - headline benchmark of the B200-native ulc-codec hot path.

Workload at N=1 (BASELINE.json configs[1], the configuration the metric is quoted on):
    batched decode-only, 4096 x 60 s stereo 44.1 kHz streams at 128 kbps CBR, BlockSize 2048, PCM16 out.
A "step" is one pass of the decoder over all streams of the rank.  Streams shard across ranks with no
data-path collective; per-GPU work is fixed as N grows ("weak").

Arms
    (default)          the CUDA path through the C ABI (libulcb200.so); `value` = inputs resident in HBM,
                       `e2e` = same metric through ULCB_DecodeStreams with HOST buffers (H2D + D2H timed).
    --impl reference   the reference's own CPU decoder (oracle/_ref/ulcdecodetool_avx, unmodified reference
                       compiled by oracle/Makefile) on all host cores; falls back to the oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE, NCH, N = 44100, 2, 2048
STREAM_SECONDS = 60
BLOCKS_PER_STREAM = (STREAM_SECONDS * RATE + N - 1) // N + 2        # 1294, tools/ulcEncodeTool.c:98
METRIC = "audio-seconds/sec (xrealtime) decode, batched 60 s stereo 44.1 kHz streams @128 kbps CBR, BlockSize 2048, PCM16 out"
POOL_FIXTURE = "music_cbr128_bs2048"


# ------------------------------------------------------------------------------------------------
# input construction: real reference-encoded blocks (tests/golden), recombined per stream
# ------------------------------------------------------------------------------------------------
def load_pool():
    from tests.fixtures import load_golden_stream
    g = load_golden_stream(POOL_FIXTURE)
    offs = np.concatenate([[0], np.cumsum(g["block_bytes"])]).astype(np.int64)
    blocks = [g["data"][offs[i]:offs[i + 1]] for i in range(2, len(offs) - 1)]   # skip the two lead-in blocks
    return blocks


def build_stream(blocks, stream_id: int, n_blocks: int = BLOCKS_PER_STREAM) -> np.ndarray:
    """Any sequence of valid blocks is a valid stream (blocks are self-delimiting, FormatSpecs.md), so a
    stream is n_blocks blocks drawn from the pool by a per-stream hash.  ~741 B/block like a real CBR-128 file."""
    rng = np.random.default_rng(0x9E3779B9 * (stream_id + 1) & 0xFFFFFFFF)
    idx = rng.integers(0, len(blocks), n_blocks)
    return np.concatenate([blocks[i] for i in idx])


def ulc_container(stream: np.ndarray, n_blocks: int) -> bytes:
    """24-byte .ulc header + payload (tools/ulc_Helper.h:10-20)."""
    hdr = struct.pack("<IHHIIHHI", 0x32434C55, N, 4096, n_blocks, RATE, NCH, 128, 24)
    return hdr + stream.tobytes()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ------------------------------------------------------------------------------------------------
def cpu_reference_decode(blocks, target_seconds: float = 12.0):
    """Times the reference's own decoder on all host cores, one process per core over disjoint streams.
    Returns dict(value=xrealtime, cores, kind, sample, wall_s)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    tool = os.path.join(ROOT, "oracle", "_ref", "ulcdecodetool_avx")
    if os.access(tool, os.X_OK):
        # ~500x realtime per core => STREAMS_PER_FILE 60 s streams take about target_seconds on one core
        per_file = max(1, int(target_seconds * 450 / STREAM_SECONDS))
        shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        with tempfile.TemporaryDirectory(dir=shm) as td:
            nb = BLOCKS_PER_STREAM * per_file
            files = []
            base = [build_stream(blocks, 100000 + i) for i in range(min(cores, 8))]
            for c in range(cores):
                s = np.concatenate([base[(c + k) % len(base)] for k in range(per_file)])
                path = os.path.join(td, f"s{c}.ulc")
                with open(path, "wb") as f:
                    f.write(ulc_container(s, nb))
                files.append(path)
            t0 = time.perf_counter()
            procs = [subprocess.Popen([tool, p, "/dev/null"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for p in files]
            rcs = [p.wait() for p in procs]
            wall = time.perf_counter() - t0
        if any(rcs):
            raise RuntimeError(f"reference decode tool failed: {rcs}")
        audio = cores * nb * N / RATE
        return {"value": audio / wall, "unit": "xrealtime", "cores": cores, "kind": "reference", "wall_s": wall,
                "sample": f"{cores} procs x {per_file} x 60 s streams, oracle/_ref/ulcdecodetool_avx (unmodified reference, AVX/FMA build), PCM16 to /dev/null"}
    # fallback: the oracle port, single core
    from oracle import oracle
    s = build_stream(blocks, 100000)
    nb = 300
    t0 = time.perf_counter()
    oracle.Decoder(NCH, N).decode_stream(s, nb)
    wall = time.perf_counter() - t0
    return {"value": nb * N / RATE / wall, "unit": "xrealtime", "cores": 1, "kind": "port", "wall_s": wall,
            "sample": f"{nb} blocks of one stream, oracle/ulc_oracle.c via ctypes (float out, no PCM16 conversion)"}


def wav_bytes(pcm16: np.ndarray) -> bytes:
    """Minimal 16-bit PCM RIFF/WAVE file (what tools/WavIO_Reader.c parses)."""
    data = np.ascontiguousarray(pcm16, np.int16).tobytes()
    nch = pcm16.shape[1]
    return (b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, nch, RATE, RATE * nch * 2, nch * 2, 16)
            + b"data" + struct.pack("<I", len(data)) + data)


def cpu_reference_encode(quality: float = 50.0, seconds: float = 30.0):
    """Times the reference's own VBR encoder (oracle/_ref/ulcencodetool_avx) on all host cores, one process per core,
    each on its own synthetic stream.  Returns dict(value=xrealtime, cores, kind, sample) or None when the tool is absent."""
    from ulc_codec_b200 import corpus
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    tool = os.path.join(ROOT, "oracle", "_ref", "ulcencodetool_avx")
    if not os.access(tool, os.X_OK):
        return None
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    ns = int(seconds * RATE)
    with tempfile.TemporaryDirectory(dir=shm) as td:
        base = [corpus.generate(200000 + i, ns, "music") for i in range(min(cores, 4))]
        files = []
        for c in range(cores):
            path = os.path.join(td, f"s{c}.wav")
            with open(path, "wb") as f:
                f.write(wav_bytes(base[c % len(base)]))
            files.append(path)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([tool, p, p + ".ulc", f"-{quality:g}", f"-blocksize:{N}"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for p in files]
        rcs = [p.wait() for p in procs]
        wall = time.perf_counter() - t0
    if any(rcs):
        return None
    return {"value": cores * seconds / wall, "unit": "xrealtime", "cores": cores, "kind": "reference",
            "sample": f"{cores} procs x one {seconds:g} s stream, oracle/_ref/ulcencodetool_avx -{quality:g} (unmodified reference, AVX/FMA build), WAV in /dev/shm"}


def verify_encode(torch, ctx, capi, rank, n_streams, ns, nb, out, ostride, bb, n_verify: int = 16, quality: float = 50.0):
    """Compares a strided sample of the encoded streams (first, last, spread over the launch) with the oracle encoder on the same
    PCM: block sizes and bytes.  BASELINE.json allows a quantisation-boundary flip in at most 0.1 % of blocks."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    from ulc_codec_b200 import corpus
    slots = sorted({0, n_streams - 1, *np.linspace(0, n_streams - 1, n_verify).astype(int).tolist()})

    def ref(s):
        pcm = corpus.generate(rank * n_streams + s, ns, "music")
        return s, oracle.Encoder(RATE, NCH, N).encode_stream(pcm, "vbr", quality, 0.0)
    with ThreadPoolExecutor(min(32, os.cpu_count() or 1)) as ex:
        refs = dict(ex.map(ref, slots))
    bbh = bb.view(n_streams, nb)
    diff = tot = 0
    for s in slots:
        r = refs[s]
        sizes = bbh[s].cpu().numpy().astype(np.int64)
        got = out[s * ostride:s * ostride + int(sizes.sum())].cpu().numpy()
        rs = r["block_bytes"].astype(np.int64)
        go = np.concatenate([[0], np.cumsum(sizes)]); ro = np.concatenate([[0], np.cumsum(rs)])
        for b in range(nb):
            tot += 1
            if sizes[b] != rs[b] or not np.array_equal(got[go[b]:go[b + 1]], r["data"][ro[b]:ro[b + 1]]):
                diff += 1
    if diff > 1e-3 * tot + 1:
        raise AssertionError(f"bench verification: {diff} of {tot} encoded blocks differ from the oracle")
    return {"streams": len(slots), "blocks": tot, "blocks_differing": diff, "against": "oracle.Encoder (byte comparison per block)"}


def bench_encode(ctx, capi, torch, dist, dev, rank, world, n_streams, steps, warmup, peaks, with_cpu):
    """configs[2]: batched VBR Quality 50 encode, 60 s stereo streams, BlockSize 2048, PCM16 in, corpus generated on the device."""
    ns = STREAM_SECONDS * RATE
    nb = ctx.blocks_for(ns)
    stride = nb * N * NCH                                           # whole blocks: the lead-out past the samples is silence
    pcm = torch.zeros(n_streams * stride, dtype=torch.int16, device=dev)
    ctx.generate_corpus_device("music", rank * n_streams, n_streams, ns, pcm.data_ptr(), stride)
    ostride = nb * 640                                               # VBR Q50 averages ~460 B/block on this corpus
    out = torch.empty(n_streams * ostride, dtype=torch.uint8, device=dev)
    bb = torch.zeros(n_streams * nb, dtype=torch.int32, device=dev)
    st = torch.zeros(n_streams, dtype=torch.int32, device=dev)
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)

    def step():
        ctx.encode_streams_device(n_streams, pcm.data_ptr(), "pcm16", stride, nb, "vbr", 50.0, out.data_ptr(), ostride,
                                  d_block_bytes=bb.data_ptr(), meta_stride=nb, d_status=st.data_ptr())
    for _ in range(warmup):
        step()
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ctx.kernel_times(reset=True)
    ctx.set_kernel_timing(True)
    e0.record(ext)
    for _ in range(steps):
        step()
    e1.record(ext)
    ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    kt = ctx.kernel_times(reset=True)
    ctx.set_kernel_timing(False)
    kernels = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for k, v in kt.items() if v[1] and k.startswith("enc_")}
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert int((st != 0).sum().item()) == 0, "encode arena overflow"
    out_bytes = int(bb.sum().item())
    algo = n_streams * nb * N * NCH * 2 + out_bytes                  # PCM16 in + bytes out
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = algo / (ms / 1e3) / 1e9
    r = {"metric": "audio-seconds/sec (xrealtime) encode, batched VBR Quality=50, 60 s stereo 44.1 kHz streams, BlockSize 2048, PCM16 in",
         "value": world * n_streams * nb * N / RATE / (ms / 1e3), "unit": "xrealtime", "ms_per_step": ms, "steps": steps, "warmup": warmup,
         "config": {"workload": "configs[2]: batched VBR encode Quality=50, window switching on", "streams_per_gpu": n_streams,
                    "blocks_per_stream": nb, "input": "integer-only synthetic music corpus generated on the device (csrc/ulcb_corpus.cu), every stream distinct",
                    "avg_bytes_per_block": out_bytes / (n_streams * nb)},
         "roofline": {"bound": "hbm", "kernel": "ulcb_detect_kernel<2,false> (transient-detector pre-pass) + ulcb_encode_kernel<2,512,false,true>",
                      "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                      "algorithmic_bytes_per_step": algo, "step_ms": ms, "kernels": kernels, "traffic": measured_traffic("encode", n_streams, nb)},
         "verified_streams": verify_encode(torch, ctx, capi, rank, n_streams, ns, nb, out, ostride, bb),
         "gpu_launches": int(sum(v["launches_per_step"] for v in kernels.values()) * steps) or steps}
    if with_cpu:
        c = cpu_reference_encode()
        r["cpu_baseline"] = c
    del pcm, out
    return r


def _event_ms(torch, ctx, ext, fn):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ctx.sync(); torch.cuda.synchronize()
    e0.record(ext); fn(); e1.record(ext)
    ctx.sync(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def _snr_db(torch, ref_i16, got_i16):
    """Per-stream SNR (dB) of the decoded PCM16 against the original, on the device."""
    r = ref_i16.to(torch.float32); e = got_i16.to(torch.float32) - r
    return 10.0 * torch.log10((r * r).sum(dim=1) / (e * e).sum(dim=1).clamp_min(1e-9))


def encode_decode_case(capi, torch, dev, rank, local, peak, *, block_size, n_streams, seconds, kind, mode, rate, first_id, verify=2, waves=1):
    """Encode n_streams x seconds of the `kind` corpus and decode the result again, the bitstream never leaving the device
    (the decoder is handed the encoder's block-size index); `waves` batches back to back reuse the same buffers.  Returns
    encode / decode / round-trip throughput, roofline fractions against the algorithmic bytes, a reconstruction SNR and a parity
    check of `verify` streams against the oracle (bytes identical, PCM16 within 1 LSB)."""
    from oracle import oracle
    from ulc_codec_b200 import corpus
    bs = block_size
    ns = seconds * RATE
    with capi.Context(RATE, NCH, bs, max_blocks=(ns + bs - 1) // bs + 2, device=local) as ctx:
        ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
        nb = ctx.blocks_for(ns)
        istride = nb * bs * NCH
        pcm = torch.zeros(n_streams * istride, dtype=torch.int16, device=dev)
        ostride = (nb * (3 * int(rate * 1000 / 8 * bs / RATE) + 64) if mode != "vbr" else nb * bs) + 64
        out = torch.zeros(n_streams * ostride + 64, dtype=torch.uint8, device=dev)
        bb = torch.zeros(n_streams * nb, dtype=torch.int32, device=dev)
        est = torch.zeros(n_streams, dtype=torch.int32, device=dev)
        offs = (torch.arange(n_streams + 1, dtype=torch.int64, device=dev) * ostride)
        dec = torch.empty(n_streams * istride, dtype=torch.int16, device=dev)
        dst = torch.zeros(n_streams, dtype=torch.int32, device=dev)

        def enc():
            ctx.encode_streams_device(n_streams, pcm.data_ptr(), "pcm16", istride, nb, mode, rate, out.data_ptr(), ostride,
                                      d_block_bytes=bb.data_ptr(), meta_stride=nb, d_status=est.data_ptr())

        def dcd():
            ctx.decode_streams_device(n_streams, out.data_ptr(), offs.data_ptr(), nb, dec.data_ptr(), istride, "pcm16",
                                      d_status=dst.data_ptr(), d_block_bytes=bb.data_ptr(), block_bytes_stride=nb)
        t_enc = t_dec = 0.0
        out_bytes = 0
        snr_min = float("inf")
        for w in range(waves + 1):                                   # wave 0 is the warm-up
            ctx.generate_corpus_device(kind, first_id + w * n_streams, n_streams, ns, pcm.data_ptr(), istride)
            te = _event_ms(torch, ctx, ext, enc)
            td = _event_ms(torch, ctx, ext, dcd)
            assert int((est != 0).sum().item()) == 0, "encode arena overflow"
            assert int((dst != -1).sum().item()) == 0, "decoder rejected an encoded stream"
            if w:
                t_enc += te; t_dec += td
                out_bytes += int(bb.sum().item())
                # the decoder's output lags the input by two blocks (encoder lap + decoder lap): compare the aligned parts
                k = min(32, n_streams)                               # a sample of streams: the float copies are 4x the PCM16
                a = pcm.view(n_streams, nb, bs * NCH)[:k, :-2].reshape(k, -1)
                b = dec.view(n_streams, nb, bs * NCH)[:k, 2:].reshape(k, -1)
                snr_min = min(snr_min, float(_snr_db(torch, a, b).min().item()))
        blocks = waves * n_streams * nb
        audio = waves * n_streams * nb * bs / RATE
        in_bytes = blocks * bs * NCH * 2
        ver = None
        if verify:
            bbh = bb.view(n_streams, nb)
            ok = 0
            for s in sorted({0, n_streams - 1, *np.linspace(0, n_streams - 1, verify).astype(int).tolist()}):
                p = corpus.generate(first_id + waves * n_streams + s, ns, kind)
                if mode == "abr2":
                    avg = float(ctx.avg_complexity(n_streams)[s])
                    ref = oracle.Encoder(RATE, NCH, bs).encode_stream(p, "abr", rate, avg)
                else:
                    ref = oracle.Encoder(RATE, NCH, bs).encode_stream(p, mode, rate, 0.0)
                sizes = bbh[s].cpu().numpy().astype(np.int64)
                got = out[s * ostride:s * ostride + int(sizes.sum())].cpu().numpy()
                nd = int((sizes != ref["block_bytes"].astype(np.int64)).sum())
                same = nd == 0 and np.array_equal(got, ref["data"])
                f, bits = oracle.Decoder(NCH, bs).decode_stream(got, nb)
                d = int(np.abs(dec[s * istride:(s + 1) * istride].cpu().numpy().reshape(-1, NCH).astype(np.int32) - oracle.float_to_pcm16(f).astype(np.int32)).max())
                if d > 1 or nd > max(1, nb // 500):
                    raise AssertionError(f"bench verification (bs {bs} {mode}): stream {s}: {nd} block sizes differ from the oracle, decode differs by {d} LSB")
                ok += 1 if same else 0
            ver = {"streams": verify, "byte_identical_to_oracle": ok, "decode_max_abs_diff_lsb": "<=1"}
        del pcm, out, dec
        torch.cuda.empty_cache()
    r = {"streams_per_gpu": n_streams, "waves": waves, "seconds_per_stream": seconds, "block_size": bs, "blocks_per_stream": nb,
         "avg_bytes_per_block": out_bytes / blocks,
         "encode": {"value": audio / (t_enc / 1e3), "unit": "xrealtime", "ms": t_enc,
                    "roofline_frac": (in_bytes + out_bytes) / (t_enc / 1e3) / 1e9 / peak},
         "decode": {"value": audio / (t_dec / 1e3), "unit": "xrealtime", "ms": t_dec,
                    "roofline_frac": (in_bytes + out_bytes) / (t_dec / 1e3) / 1e9 / peak},
         "round_trip": {"value": audio / ((t_enc + t_dec) / 1e3), "unit": "xrealtime"},
         "min_snr_db": snr_min, "verification": ver}
    return r


def cpu_reference_tool(kind, args, seconds=20.0, decode=False):
    """The reference tool (AVX build) on all host cores, one process per core, for an arbitrary tool command line."""
    from ulc_codec_b200 import corpus
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    enc_tool = os.path.join(ROOT, "oracle", "_ref", "ulcencodetool_avx")
    dec_tool = os.path.join(ROOT, "oracle", "_ref", "ulcdecodetool_avx")
    if not (os.access(enc_tool, os.X_OK) and os.access(dec_tool, os.X_OK)):
        return None
    shm = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    ns = int(seconds * RATE)
    with tempfile.TemporaryDirectory(dir=shm) as td:
        base = [corpus.generate(300000 + i, ns, kind) for i in range(min(cores, 4))]
        files = []
        for c in range(cores):
            path = os.path.join(td, f"s{c}.wav")
            with open(path, "wb") as f:
                f.write(wav_bytes(base[c % len(base)]))
            files.append(path)
        t0 = time.perf_counter()
        procs = [subprocess.Popen([enc_tool, p, p + ".ulc", *args], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for p in files]
        rcs = [p.wait() for p in procs]
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        procs = [subprocess.Popen([dec_tool, p + ".ulc", "/dev/null"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for p in files]
        rcs += [p.wait() for p in procs]
        t_dec = time.perf_counter() - t0
    if any(rcs):
        return None
    return {"encode": cores * seconds / t_enc, "decode": cores * seconds / t_dec, "unit": "xrealtime", "cores": cores, "kind": "reference",
            "sample": f"{cores} procs x one {seconds:g} s {kind} stream, oracle/_ref/ulc{{en,de}}codetool_avx {' '.join(args)}"}


# ------------------------------------------------------------------------------------------------

def build_id() -> str:
    """Hash of the CUDA sources the loaded library was built from (profiles/traffic.json is keyed by it)."""
    import glob
    import hashlib
    h = hashlib.sha1()
    for f in sorted(glob.glob(os.path.join(ROOT, "ulc_codec_b200", "csrc", "*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:12]


def measured_traffic(kernel: str, n_streams: int, n_blocks: int):
    """DRAM bytes (read + write) per step of `kernel`, from the committed `ncu --set full` capture of the same launch shape AND the
    same build (profiles/traffic.json, written by scripts/ncu_traffic.py).  None when shape or build differ."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
    except Exception:
        return None
    if t.get("streams") != n_streams or t.get("blocks") != n_blocks or t.get("build") != build_id():
        return None
    return int(t["dram_bytes_read"] + t["dram_bytes_write"])


def make_streams(pool, n_distinct: int, seed0: int, n_blocks: int = BLOCKS_PER_STREAM):
    """n_distinct streams of n_blocks pool blocks each.  Returns (streams, block-size index uint32[n_distinct, n_blocks])."""
    sizes = np.array([len(b) for b in pool], np.uint32)
    streams, idx = [], np.zeros((n_distinct, n_blocks), np.uint32)
    for d in range(n_distinct):
        rng = np.random.default_rng(0x9E3779B9 * (seed0 + d + 1) & 0xFFFFFFFF)
        sel = rng.integers(0, len(pool), n_blocks)
        streams.append(np.concatenate([pool[i] for i in sel]))
        idx[d] = sizes[sel]
    return streams, idx


class DecodeBatch:
    """A device-resident decode batch: n_streams slots filled from n_distinct distinct streams in an order in which neighbours
    (and slots one launch wave apart) differ."""

    def __init__(self, torch, dev, distinct, index, n_streams: int):
        self.torch, self.n_streams, self.n_distinct = torch, n_streams, len(distinct)
        self.distinct, self.index = distinct, index
        nd = self.n_distinct
        self.order = [(s * 7 + s // nd) % nd for s in range(n_streams)]
        lens = np.array([len(distinct[o]) for o in self.order], np.uint64)
        self.offs = np.zeros(n_streams + 1, np.uint64)
        self.offs[1:] = np.cumsum(lens)
        self.total_in = int(self.offs[-1])
        d_dist = [torch.from_numpy(d).to(dev) for d in distinct]
        self.d_bytes = torch.empty(self.total_in + 64, dtype=torch.uint8, device=dev)
        for s, o in enumerate(self.order):
            self.d_bytes[int(self.offs[s]):int(self.offs[s + 1])] = d_dist[o]
        self.d_bytes[self.total_in:] = 0
        del d_dist
        self.d_offs = torch.from_numpy(self.offs.view(np.int64)).to(dev)
        self.d_index = torch.from_numpy(index.view(np.int32)).to(dev)[torch.tensor(self.order, device=dev)].contiguous()
        self.stride = BLOCKS_PER_STREAM * N * NCH
        self.d_status = torch.zeros(n_streams, dtype=torch.int32, device=dev)


def time_steps(torch, dist, ctx, ext, dev, world, step, steps: int, clock_gpu=None):
    """Times `steps` calls of step() with CUDA events on the library's stream; returns (ms per step over the whole region as
    the max over ranks, mean ms of the individual steps, clocks summary or None)."""
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(clock_gpu) if clock_gpu is not None else None
    if clk:
        clk.__enter__()
    t0.record(ext)
    for a, b in evs:
        a.record(ext); step(); b.record(ext)
    t1.record(ext)
    ctx.sync()
    torch.cuda.synchronize()
    if clk:
        clk.__exit__()
    total = t0.elapsed_time(t1)
    each = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    if world > 1:
        t = torch.tensor([total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
        dist.barrier()
    return total / steps, each, (clk.summary() if clk else None)


def verify_decode(torch, batch: DecodeBatch, d_out, n_verify: int = 64):
    """Copies back a strided sample of the decoded streams (first, last, and slots spread over every launch wave) and compares each
    with the oracle decoder: PCM16 within +-1 LSB, whole stream.  Returns the number of streams verified (raises on a mismatch)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    ns = batch.n_streams
    slots = sorted({0, ns - 1, *np.linspace(0, ns - 1, n_verify).astype(int).tolist()})
    need = sorted({batch.order[s] for s in slots})

    def ref(o):
        f, bits = oracle.Decoder(NCH, N).decode_stream(batch.distinct[o], BLOCKS_PER_STREAM)
        assert (bits != 0).all()
        return o, oracle.float_to_pcm16(f)
    with ThreadPoolExecutor(min(32, os.cpu_count() or 1)) as ex:
        refs = dict(ex.map(ref, need))
    worst = 0
    for s in slots:
        got = d_out[s * batch.stride:(s + 1) * batch.stride].cpu().numpy().reshape(-1, NCH)
        d = int(np.abs(got.astype(np.int32) - refs[batch.order[s]].astype(np.int32)).max())
        if d > 1:
            raise AssertionError(f"bench verification: stream slot {s} differs from the oracle by {d} LSB")
        worst = max(worst, d)
    return len(slots), worst


def pcie_ceiling(torch, dist, dev, world, gb: float = 2.0):
    """What the box gives this rank for plain pinned-memory copies while every rank copies at once (cudaMemcpyAsync, GB/s)."""
    n = int(gb * (1 << 30))
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    res = {}
    for name, (dst, src) in {"d2h_gbs": (h, d), "h2d_gbs": (d, h)}.items():
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        if world > 1:
            t = torch.tensor([dt], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
        res[name] = n / dt / 1e9
    del h, d
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU (default = BASELINE configs[1])")
    ap.add_argument("--distinct", type=int, default=1024, help="distinct streams behind the batch")
    ap.add_argument("--e2e-streams", type=int, default=4096, help="streams per GPU of the host-buffer (e2e) measurement")
    ap.add_argument("--verify", type=int, default=64, help="decoded streams compared with the oracle after the timed loop (0 = skip)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--enc-streams", type=int, default=4096, help="streams per GPU of the secondary encode measurement (0 = skip)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary objects (indexed / mixed decode, configs[3], configs[4])")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = {"workload": "configs[1]: batched decode-only, 60 s stereo streams @128 kbps CBR, BlockSize 2048, PCM16 out",
           "streams_per_gpu": args.streams, "blocks_per_stream": BLOCKS_PER_STREAM, "rate_hz": RATE, "n_chan": NCH, "block_size": N,
           "input": f"reference-encoded (oracle/_ref scalar build) CBR-128 blocks of the synthetic music corpus; each stream = 1294 blocks "
                    f"drawn per stream from the 431-block golden pool tests/golden/music_cbr128_bs2048.npz; {min(args.distinct, args.streams)} "
                    "distinct streams per GPU, replicated; raw .ulc payloads WITHOUT a block index (what the reference decoder reads), "
                    "so every step includes the token-length walk that locates the blocks",
           "l2": "inputs+outputs per step (>40 GB) far exceed the 126 MB L2; no flush needed"}
    blocks = load_pool()

    if args.impl == "reference":
        if rank != 0:
            return
        vals = []
        for i in range(args.warmup + args.steps):
            r = cpu_reference_decode(blocks, target_seconds=4.0)
            if i >= args.warmup:
                vals.append(r)
        wall = float(np.mean([v["wall_s"] for v in vals]))
        value = float(np.mean([v["value"] for v in vals]))
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "xrealtime", "n_gpus": 0, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": value, "unit": "xrealtime", "cores": vals[-1]["cores"], "kind": vals[-1]["kind"], "sample": vals[-1]["sample"]},
                "e2e": {"value": value, "unit": "xrealtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from ulc_codec_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"

    # ---- inputs (untimed) ----
    n_streams = args.streams
    distinct, index = make_streams(blocks, min(args.distinct, n_streams), rank * 100003)
    batch = DecodeBatch(torch, dev, distinct, index, n_streams)
    stride = batch.stride
    d_out = torch.empty(n_streams * stride, dtype=torch.int16, device=dev)

    ctx = capi.Context(RATE, NCH, N, max_blocks=BLOCKS_PER_STREAM, device=local)
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)

    def step(indexed=False, b=batch):
        kw = dict(d_block_bytes=b.d_index.data_ptr(), block_bytes_stride=BLOCKS_PER_STREAM) if indexed else {}
        ctx.decode_streams_device(b.n_streams, b.d_bytes.data_ptr(), b.d_offs.data_ptr(), BLOCKS_PER_STREAM, d_out.data_ptr(), stride,
                                  "pcm16", d_status=b.d_status.data_ptr(), **kw)

    # ---- headline: foreign streams, inputs resident ----
    for _ in range(args.warmup):
        step()
    ctx.set_kernel_timing(True)
    ctx.kernel_times(reset=True)
    l0 = ctx.launches
    ms_per_step, kern_ms, clocks = time_steps(torch, dist, ctx, ext, dev, world, step, args.steps, clock_gpu=local)
    launches = ctx.launches - l0
    ktimes = ctx.kernel_times(reset=True)
    pipe = ctx.decode_pipeline()
    ktimes_alone = None
    if pipe[1] > 0:
        # the timed region ran the chunked pipeline (kernels of neighbouring chunks side by side on two streams, so their spans overlap
        # and their sum exceeds the step); the same kernels one at a time, for the per-kernel durations the ncu launch list can be held to
        ctx.set_decode_pipeline(1, 0)
        step()
        ctx.kernel_times(reset=True)
        for _ in range(args.steps):
            step()
        ktimes_alone = ctx.kernel_times(reset=True)
        ctx.set_decode_pipeline(*pipe)
    ctx.set_kernel_timing(False)
    assert int((batch.d_status != -1).sum().item()) == 0, "decode reported corrupt streams"
    audio_per_step = world * n_streams * BLOCKS_PER_STREAM * N / RATE
    value = audio_per_step / (ms_per_step / 1e3)
    verified = None
    if args.verify > 0:
        nv, worst = verify_decode(torch, batch, d_out, args.verify)
        verified = {"streams": nv, "max_abs_diff_lsb": worst, "against": "oracle.Decoder (pinned to the reference's scalar build), whole 60 s streams, PCM16"}

    # ---- roofline of the dominant kernel (the transform: ~3/4 of the step): algorithmic bytes / its time per step ----
    algo_bytes = batch.total_in + n_streams * stride * 2            # bytes in + PCM16 out (8934 B per block)
    per_kernel = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in ktimes.items() if v[1]}
    tr_ms = per_kernel.get("bp_transform", {}).get("ms_per_step") or kern_ms
    achieved = algo_bytes / (tr_ms / 1e3) / 1e9
    roof = {"bound": "hbm", "kernel": "bp_transform_kernel<true,false>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_source, "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": tr_ms,
            "launch_note": "one step = the pipeline walk -> (parse -> scan -> transform) per wave of streams; launch_ms is the transform "
                           "kernel's summed time per step (CUDA events on the library's stream around every launch), algorithmic bytes are "
                           "the whole step's (bytes in + PCM16 out)",
            "kernels_ms_per_step": per_kernel,
            **({"pipeline": {"chunks": pipe[0], "overlap": pipe[1],
                             "note": "ULCB_SetDecodePipeline: upload / walk (overlap 1) and parse (overlap 2) of chunk c+1 on side streams beside the "
                                     "transform of chunk c; kernels_ms_per_step are the spans inside the timed region (overlapping), "
                                     "kernels_ms_per_step_alone the same kernels run one at a time (pipeline off, extra untimed pass)"},
                "kernels_ms_per_step_alone": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                                              for k, v in ktimes_alone.items() if v[1]}} if ktimes_alone else {}),
            "step": {"ms": kern_ms, "achieved": algo_bytes / (kern_ms / 1e3) / 1e9, "frac": algo_bytes / (kern_ms / 1e3) / 1e9 / peak},
            "traffic": measured_traffic("decode", n_streams, BLOCKS_PER_STREAM)}

    extra = {}
    if not args.no_extra:
        # the same batch with the caller-supplied block-size index (this library's encoder and container.py emit one): no walk
        for _ in range(2):
            step(True)
        ms_i, _, _ = time_steps(torch, dist, ctx, ext, dev, world, lambda: step(True), max(2, args.steps // 2))
        extra["decode_indexed"] = {"value": audio_per_step / (ms_i / 1e3), "unit": "xrealtime", "ms_per_step": ms_i,
                                   "roofline_step_frac": algo_bytes / (ms_i / 1e3) / 1e9 / peak,
                                   "note": "ULCB_DecodeStreamsIndexedDevice: block sizes supplied, no token-length walk"}
        assert int((batch.d_status != -1).sum().item()) == 0

    # ---- e2e: through the host-pointer C-ABI call, pinned host buffers, copies inside the timed region ----
    e2e = None
    pcie = None
    if rank == 0 or world > 1:
        pcie = pcie_ceiling(torch, dist, dev, world)
        ne = min(args.e2e_streams, n_streams)
        # every rank page-locks its own PCM sink (10.6 MB per stream, 43 GB for the whole 4096-stream batch): all ranks together stay
        # within a quarter of the host memory that is free now; every rank takes the same count
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable:"))
            fit = int(avail / 4 / max(1, world) / (stride * 2 + 1.0e6))
            if world > 1:
                t = torch.tensor([fit], device=dev, dtype=torch.int64); dist.all_reduce(t, op=dist.ReduceOp.MIN); fit = int(t.item())
            ne = max(64, min(ne, fit))
        except (OSError, StopIteration, ValueError):
            pass
        h_streams = [distinct[o] for o in batch.order[:ne]]
        h_data, h_offs = capi.pack_streams(h_streams)
        pin_in = torch.empty(len(h_data), dtype=torch.uint8).pin_memory()
        pin_in.numpy()[:] = h_data
        pin_out = torch.empty(ne * stride, dtype=torch.int16).pin_memory()
        h_status = np.zeros(ne, np.int32)

        def e2e_step(n=ne, out=pin_out):
            rc = ctx.L.ULCB_DecodeStreams(ctx.h, n, pin_in.data_ptr(), h_offs.ctypes.data, None, BLOCKS_PER_STREAM,
                                          out.data_ptr(), stride, capi.FMT_PCM16, h_status.ctypes.data, None, 0)
            assert rc == 0, rc
        e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        reps = max(2, args.steps // 2)
        for _ in range(reps):
            e2e_step()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert (h_status == -1).all()
        if args.verify > 0:                                            # the host buffer holds what the device path produced
            # streams spread over the chunks of the pipelined call (34 of them, first and last included): bit-identical to the
            # device-resident run, whose streams verify_decode() holds to the oracle
            chk = sorted({0, ne - 1, *(int(v) for v in np.linspace(0, ne - 1, 34))})
            for s_chk in chk:
                ref = d_out[s_chk * stride:(s_chk + 1) * stride].cpu().numpy()
                assert np.array_equal(pin_out.numpy()[s_chk * stride:(s_chk + 1) * stride], ref), f"e2e output of stream {s_chk} differs from the device-resident run"
        d2h = int(ne * stride * 2)
        e2e = {"value": world * ne * BLOCKS_PER_STREAM * N / RATE / dt, "unit": "xrealtime", "h2d_bytes_per_step": int(len(h_data)),
               "d2h_bytes_per_step": d2h, "streams_per_gpu": ne, "ms_per_step": dt * 1e3,
               "verified_streams": (len(chk) if args.verify > 0 else 0),
               "pcie_gbs": (len(h_data) + d2h) / dt / 1e9, "pcie_ceiling": pcie,
               "api": "ULCB_DecodeStreams (host pointers, pinned; bytes in by H2D copy, PCM out stored by the kernel straight into the pinned "
                      "buffer over PCIe); streams_per_gpu = the whole batch unless a quarter of the free host memory, shared by the ranks, holds fewer",
               "pipeline": "library default for host-pointer calls (ULCB_SetDecodePipeline, chunks 0 = automatic): chunks of 128 streams, the upload, "
                           "token-length walk and parse of chunk c+1 on side CUDA streams beside the transform of chunk c, whose stores wait for PCIe"}
        if not args.no_extra and rank == 0 and world == 1:
            # pageable caller buffers: staged through device memory and one D2H copy
            npg = min(256, ne)
            pg_out = torch.empty(npg * stride, dtype=torch.int16)
            e2e_step(npg, pg_out)
            t0 = time.perf_counter()
            e2e_step(npg, pg_out)
            dtp = time.perf_counter() - t0
            e2e["pageable"] = {"value": npg * BLOCKS_PER_STREAM * N / RATE / dtp, "unit": "xrealtime", "streams": npg, "ms_per_step": dtp * 1e3}
            del pg_out
        del pin_in, pin_out

    if not args.no_extra:
        # window-switched blocks priced in: the same launch shape on a pool that mixes the music stream with the transient-heavy ones
        from tests import wavecheck
        mixed_pool = wavecheck.block_pool()
        nm = min(1024, n_streams)
        md, mi = make_streams(mixed_pool, min(256, nm), 7000 + rank * 100003)
        mb = DecodeBatch(torch, dev, md, mi, nm)
        for _ in range(2):
            step(False, mb)
        ms_m, _, _ = time_steps(torch, dist, ctx, ext, dev, world, lambda: step(False, mb), max(2, args.steps // 2))
        assert int((mb.d_status != -1).sum().item()) == 0
        extra["decode_mixed"] = {"value": world * nm * BLOCKS_PER_STREAM * N / RATE / (ms_m / 1e3), "unit": "xrealtime", "ms_per_step": ms_m,
                                 "streams_per_gpu": nm, "avg_bytes_per_block": mb.total_in / (nm * BLOCKS_PER_STREAM),
                                 "note": "pool = music_cbr128 + transient_vbr50 + music_vbr75 goldens (about a third of the blocks window-switched "
                                         "or transient-flagged), no block index"}
        del mb

    # ---- secondary measurement: batched VBR encode (configs[2]); the headline stays the decode line above ----
    enc = None
    del batch.d_bytes, d_out
    ctx.close()                                                      # gives the decode scratch (sparse-content arena) back
    torch.cuda.empty_cache()
    ctx = capi.Context(RATE, NCH, N, max_blocks=BLOCKS_PER_STREAM, device=local)
    if args.enc_streams > 0:
        enc = bench_encode(ctx, capi, torch, dist, dev, rank, world, args.enc_streams, max(1, args.steps // 2), 1, peaks,
                           with_cpu=(rank == 0 and world == 1 and not args.no_cpu))

    # ---- configs[3] (transient-heavy, BlockSize 256 / 8192, ABR-64 with the quality search on the device) and a bounded sample of
    #      configs[4] (VBR Q75 encode -> decode round trip, bitstream resident, in waves) ----
    if not args.no_extra:
        torch.cuda.empty_cache()
        c3 = {}
        for bs, nst in ((256, 1024), (8192, 512)):
            c3[f"bs{bs}"] = encode_decode_case(capi, torch, dev, rank, local, peak, block_size=bs, n_streams=nst, seconds=60, kind="transient",
                                               mode="abr2", rate=64.0, first_id=500000 + rank * 10000)
            if rank == 0 and world == 1 and not args.no_cpu:
                c3[f"bs{bs}"]["cpu_baseline"] = cpu_reference_tool("transient", ["64,0.5", f"-blocksize:{bs}"])
        extra["configs3_transient_abr64"] = {"workload": "configs[3]: transient-heavy corpus (impulse trains + noise bursts), ABR 64 kbps, two passes on the "
                                                         "device (analysis pass for AvgComplexity, then the size search per block), then decode of the result",
                                             **c3}
        rt = encode_decode_case(capi, torch, dev, rank, local, peak, block_size=N, n_streams=1024, seconds=180, kind="music", mode="vbr", rate=75.0,
                                first_id=700000 + rank * 100000, waves=2, verify=2)
        if world > 1:
            t = torch.tensor([rt["encode"]["ms"] + rt["decode"]["ms"]], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rt["round_trip"]["value"] = world * 2 * 1024 * rt["blocks_per_stream"] * N / RATE / (float(t.item()) / 1e3)
        if rank == 0 and world == 1 and not args.no_cpu:
            rt["cpu_baseline"] = cpu_reference_tool("music", ["-75"])
        extra["configs4_roundtrip_q75"] = {"workload": "configs[4]: VBR Quality=75 encode + decode round trip of 180 s stereo streams, stream-sharded; the full "
                                                       "corpus is 65536 streams = 8192 per GPU at 8 GPUs, processed in resident waves; this is a bounded sample "
                                                       "of 2 waves x 1024 streams per GPU (a wave of 2048 streams is 65 GB of PCM16 and fits as well)", **rt}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu:
        c = cpu_reference_decode(blocks)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {"metric": METRIC, "value": value, "unit": "xrealtime", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "verified_streams": verified["streams"] if verified else 0, "verification": verified,
            "roofline": roof, "cpu_baseline": cpu, "encode": enc, "build": build_id(), **extra}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
