"""GPU: time the batched decoder on golden-pool streams (device-resident input) + phase cycles."""
import sys, os, argparse, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ulc_codec_b200 import capi
ap = argparse.ArgumentParser(); ap.add_argument("--streams", type=int, default=444); ap.add_argument("--blocks", type=int, default=300); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
blocks = bench.load_pool(); dev = torch.device("cuda", 0)
streams = [bench.build_stream(blocks, d, a.blocks) for d in range(min(64, a.streams))]
order = [s % len(streams) for s in range(a.streams)]
data, offs = capi.pack_streams([streams[o] for o in order])
d_bytes = torch.from_numpy(np.concatenate([data, np.zeros(64, np.uint8)])).to(dev); d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
stride = a.blocks * 2048 * 2
d_out = torch.empty(a.streams * stride, dtype=torch.int16, device=dev); d_st = torch.zeros(a.streams, dtype=torch.int32, device=dev)
with capi.Context(44100, 2, 2048, max_blocks=a.blocks) as ctx:
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
    step = lambda: ctx.decode_streams_device(a.streams, d_bytes.data_ptr(), d_offs.data_ptr(), a.blocks, d_out.data_ptr(), stride, "pcm16", d_status=d_st.data_ptr())
    step(); ctx.sync()
    if hasattr(ctx.L, "ULCB_DebugDecodePhaseTimes"):
        buf = (C.c_ulonglong * 16)(); ctx.L.ULCB_DebugDecodePhaseTimes(buf)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(a.reps): step()
    e1.record(ext); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    nblk = a.streams * a.blocks
    print(f"decode streams={a.streams} blocks={a.blocks}: {ms:.2f} ms, {nblk * 2048 / 44100 / (ms / 1e3):.0f} x realtime, {(len(data) + nblk * 8192) / ms / 1e6:.1f} GB/s, bad {int((d_st != -1).sum())}")
    if hasattr(ctx.L, "ULCB_DebugDecodePhaseTimes"):
        buf = (C.c_ulonglong * 16)(); ctx.L.ULCB_DebugDecodePhaseTimes(buf)
        n = max(1, buf[8]); names = ["parse", "noise", "dct", "unfold+delay", "store"]
        print("phase cycles per block (thread 0): " + ", ".join(f"{nm} {buf[i] / n:.0f}" for i, nm in enumerate(names)) + f" | total {sum(buf[i] for i in range(8)) / n:.0f} | ensure/load_page {buf[5] / n:.0f} resolve {buf[6] / n:.0f} emit {buf[7] / n:.0f}")
    if hasattr(ctx.L, "ULCB_DebugDecodeWsPhaseTimes"):
        buf = (C.c_ulonglong * 32)(); ctx.L.ULCB_DebugDecodeWsPhaseTimes(buf)
        n = max(1, buf[8])
        print("ws parse cycles/block: " + ", ".join(f"{nm} {buf[i] / n:.0f}" for i, nm in enumerate(["other", "wait-empty", "win-load", "chain"])) + f" | chains/block {buf[4] / n:.2f}")
        names = ["other", "wait-full", "noise", "load+pre", "fft1+tw", "transpose", "fft2", "post+unfold", "swap-write", "xchg-wait", "ms+store"]
        print("ws transform(ch0) cycles/block: " + ", ".join(f"{nm} {buf[16 + i] / n:.0f}" for i, nm in enumerate(names)))
        print("ws chain cycles/block: " + ", ".join(f"{nm} {buf[9 + i] / n:.0f}" for i, nm in enumerate(["digest", "mapscan", "entry+walk", "counts+lq", "scans", "end-walk", "emit"])))
