"""GPU: time the batched decoder on golden-pool streams (device-resident input).
   python scripts/time_decode.py [--streams 1024] [--blocks 1294] [--reps 3] [--mode index|walk]
   Kernel choice through the environment: ULCB_DECODE_WS=1 / ULCB_DECODE_GENERIC=1 select the stream-serial kernels."""
import sys, os, argparse
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ulc_codec_b200 import capi
if os.environ.get("ULCB_LIB"): capi.LIB_PATH = os.path.abspath(os.environ["ULCB_LIB"])      # experiment builds (scripts/build_variant.sh)
ap = argparse.ArgumentParser(); ap.add_argument("--streams", type=int, default=1024); ap.add_argument("--blocks", type=int, default=1294)
ap.add_argument("--reps", type=int, default=3); ap.add_argument("--mode", default="index"); ap.add_argument("--distinct", type=int, default=64)
a = ap.parse_args()
blocks = bench.load_pool(); dev = torch.device("cuda", 0)
rng = np.random.default_rng(5)
sel = [rng.integers(0, len(blocks), a.blocks) for d in range(min(a.distinct, a.streams))]
streams = [np.concatenate([blocks[i] for i in ix]) for ix in sel]
sizes = [np.array([len(blocks[i]) for i in ix], np.uint32) for ix in sel]
order = [s % len(streams) for s in range(a.streams)]
data, offs = capi.pack_streams([streams[o] for o in order])
idx = np.stack([sizes[o] for o in order])
d_bytes = torch.from_numpy(np.concatenate([data, np.zeros(64, np.uint8)])).to(dev); d_offs = torch.from_numpy(offs.view(np.int64)).to(dev)
d_idx = torch.from_numpy(idx.view(np.int32)).to(dev)
stride = a.blocks * 2048 * 2
d_out = torch.empty(a.streams * stride, dtype=torch.int16, device=dev); d_st = torch.zeros(a.streams, dtype=torch.int32, device=dev)
with capi.Context(44100, 2, 2048, max_blocks=a.blocks) as ctx:
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
    kw = dict(d_block_bytes=d_idx.data_ptr(), block_bytes_stride=a.blocks) if a.mode == "index" else {}
    step = lambda: ctx.decode_streams_device(a.streams, d_bytes.data_ptr(), d_offs.data_ptr(), a.blocks, d_out.data_ptr(), stride, "pcm16", d_status=d_st.data_ptr(), **kw)
    step(); ctx.sync()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(a.reps): step()
    e1.record(ext); ctx.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    nblk = a.streams * a.blocks
    print(f"decode[{a.mode}{' ws' if os.environ.get('ULCB_DECODE_WS') else ''}{' generic' if os.environ.get('ULCB_DECODE_GENERIC') else ''}] streams={a.streams} blocks={a.blocks}: {ms:.2f} ms, "
          f"{nblk * 2048 / 44100 / (ms / 1e3) / 1e6:.3f} M x realtime, {(len(data) + nblk * 8192) / ms / 1e6:.1f} GB/s algorithmic, bad {int((d_st != -1).sum())}, launches/step {ctx.launches // (a.reps + 1)}")
