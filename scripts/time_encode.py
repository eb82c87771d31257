"""GPU: time the batched encoder on replicated synthetic streams (device-resident input)."""
import sys, os, time, argparse
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ulc_codec_b200 import capi, corpus
if os.environ.get("ULCB_LIB"): capi.LIB_PATH = os.path.abspath(os.environ["ULCB_LIB"])      # experiment builds (scripts/build_variant.sh)

ap = argparse.ArgumentParser()
ap.add_argument("--streams", type=int, default=296)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--bs", type=int, default=2048)
ap.add_argument("--mode", default="vbr")
ap.add_argument("--rate", type=float, default=50.0)
ap.add_argument("--distinct", type=int, default=8)
ap.add_argument("--kind", default="music")
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
N, C = a.bs, 2
ns = int(a.seconds * 44100)
dev = torch.device("cuda", 0)
with capi.Context(44100, C, N) as ctx:
    nb = ctx.blocks_for(ns)
    stride = ns * C
    base = [torch.from_numpy(corpus.generate(1000 + d, ns, a.kind).reshape(-1)).to(dev) for d in range(a.distinct)]
    pcm = torch.empty(a.streams * stride, dtype=torch.int16, device=dev)
    for s in range(a.streams):
        pcm[s * stride:(s + 1) * stride] = base[s % a.distinct]
    ostride = nb * 1200 if a.mode == "vbr" else nb * 2048
    out = torch.empty(a.streams * ostride, dtype=torch.uint8, device=dev)
    bb = torch.zeros(a.streams * nb, dtype=torch.int32, device=dev)
    st = torch.zeros(a.streams, dtype=torch.int32, device=dev)
    nsamp = torch.full((a.streams,), ns, dtype=torch.int64, device=dev)
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
    def step():
        ctx.encode_streams_device(a.streams, pcm.data_ptr(), "pcm16", stride, nb, a.mode, a.rate, out.data_ptr(), ostride,
                                  avg_complexity=0.8, d_nsamples=nsamp.data_ptr(), d_block_bytes=bb.data_ptr(), meta_stride=nb, d_status=st.data_ptr())
    torch.cuda.synchronize()
    step(); ctx.sync()
    ctx.set_kernel_timing(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(a.reps): step()
    e1.record(ext); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    audio = a.streams * nb * N / 44100
    byts = int(bb.sum().item())
    print(f"encode {a.mode} {a.rate} N={N} streams={a.streams} blocks/stream={nb}: {ms:.1f} ms/launch, {audio / (ms / 1e3):.0f} x realtime, "
          f"{ms * 1e3 / (nb * a.streams) * min(a.streams, 148):.1f} us per block per SM-slot, avg {byts / (a.streams * nb):.1f} B/block, status bad {int((st != 0).sum())}")

    kt = ctx.kernel_times()
    print("per launch: " + ", ".join(f"{k} {v[0] / max(1, a.reps):.2f} ms" for k, v in kt.items() if v[1]))
    import ctypes as C
    if hasattr(ctx.L, "ULCB_DebugPhaseTimes"):
        buf = (C.c_ulonglong * 32)()
        ctx.L.ULCB_DebugPhaseTimes(buf)
        nblk = max(1, buf[12])
        names = ["ingest", "wait-detector", "begin+fold", "dct", "analyse", "serial-sums", "spread", "keys+sort", "rate+encode", "emit"]
        tot = sum(buf[i] for i in range(10))
        print("phase cycles per block (thread 0): " + ", ".join(f"{n} {buf[i] / nblk:.0f}" for i, n in enumerate(names)) + f" | total {tot / nblk:.0f} | begin_block {buf[10] / nblk:.0f} detector busy {buf[13] / nblk:.0f} wait {buf[14] / nblk:.0f} [stage {buf[16] / nblk:.0f} energy {buf[17] / nblk:.0f} post {buf[18] / nblk:.0f} pre {buf[19] / nblk:.0f} blk {buf[20] / nblk:.0f} decide {buf[21] / nblk:.0f}]")
        print("encode_pass cycles per block: " + ", ".join(f"{n} {buf[22 + i] / nblk:.0f}" for i, n in enumerate(["select", "zones", "zone-walk", "zone-info", "kept", "walk-count", "scan", "walk-write"])))
