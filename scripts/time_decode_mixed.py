"""GPU: time the batched decoder on the mixed pool (music + transient goldens: about a third of the blocks window-switched)."""
import sys, os, argparse
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tests import wavecheck
from ulc_codec_b200 import capi
if os.environ.get("ULCB_LIB"): capi.LIB_PATH = os.path.abspath(os.environ["ULCB_LIB"])
ap = argparse.ArgumentParser(); ap.add_argument("--streams", type=int, default=1024); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--pool", default="mixed")
a = ap.parse_args()
dev = torch.device("cuda", 0)
pool = wavecheck.block_pool() if a.pool == "mixed" else wavecheck.block_pool((a.pool,))
md, mi = bench.make_streams(pool, min(256, a.streams), 7000)
b = bench.DecodeBatch(torch, dev, md, mi, a.streams)
d_out = torch.empty(a.streams * b.stride, dtype=torch.int16, device=dev)
with capi.Context(44100, 2, 2048, max_blocks=bench.BLOCKS_PER_STREAM) as ctx:
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
    step = lambda: ctx.decode_streams_device(a.streams, b.d_bytes.data_ptr(), b.d_offs.data_ptr(), bench.BLOCKS_PER_STREAM, d_out.data_ptr(), b.stride, "pcm16",
                                             d_status=b.d_status.data_ptr(), d_block_bytes=b.d_index.data_ptr(), block_bytes_stride=bench.BLOCKS_PER_STREAM)
    step(); ctx.sync()
    ctx.set_kernel_timing(True); ctx.kernel_times(reset=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(a.reps): step()
    e1.record(ext); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    kt = ctx.kernel_times()
    print(f"decode[{a.pool}, indexed] streams={a.streams}: {ms:.2f} ms, {a.streams * bench.BLOCKS_PER_STREAM * 2048 / 44100 / (ms / 1e3) / 1e6:.3f} M x realtime, bad {int((b.d_status != -1).sum())}, "
          + ", ".join(f"{k} {v[0] / a.reps:.2f}" for k, v in kt.items() if v[1]))
