"""GPU: the chunked decode pipeline (ULCB_SetDecodePipeline) against the one-chunk call, same process, same inputs.
   python scripts/time_pipeline.py [--streams 4096] [--e2e-streams 2048] [--reps 4] [--configs 1:0,4:1,8:1,4:2,8:2]
For every (chunks:overlap): device-resident foreign decode, indexed decode and the host-pointer (pinned, zero-copy PCM out) call;
the PCM of every configuration must equal the one-chunk call's bit for bit."""
import sys, os, argparse, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ulc_codec_b200 import capi
ap = argparse.ArgumentParser(); ap.add_argument("--streams", type=int, default=4096); ap.add_argument("--blocks", type=int, default=1294)
ap.add_argument("--e2e-streams", type=int, default=2048); ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--distinct", type=int, default=256); ap.add_argument("--configs", default="1:0,4:1,8:1,16:1,4:2,8:2,16:2")
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
ne = min(a.e2e_streams, a.streams)
h_data, h_offs = capi.pack_streams([streams[o] for o in order[:ne]])
pin_in = torch.empty(len(h_data), dtype=torch.uint8).pin_memory(); pin_in.numpy()[:] = h_data
pin_out = torch.empty(ne * stride, dtype=torch.int16).pin_memory() if ne else None
h_status = np.zeros(max(1, ne), np.int32)
audio = lambda n: n * a.blocks * 2048 / 44100
res = []
with capi.Context(44100, 2, 2048, max_blocks=a.blocks) as ctx:
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)

    def timed(fn, reps):
        fn(); ctx.sync()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps): fn()
        e1.record(ext); ctx.sync(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ref_sum = None
    for cfg in a.configs.split(","):
        ch, ov = (int(v) for v in cfg.split(":"))
        ctx.set_decode_pipeline(ch, ov)
        row = {"chunks": ch, "overlap": ov}
        for mode in ("walk", "index"):
            kw = dict(d_block_bytes=d_idx.data_ptr(), block_bytes_stride=a.blocks) if mode == "index" else {}
            d_out.zero_()
            step = lambda: ctx.decode_streams_device(a.streams, d_bytes.data_ptr(), d_offs.data_ptr(), a.blocks, d_out.data_ptr(), stride, "pcm16", d_status=d_st.data_ptr(), **kw)
            ms = timed(step, a.reps)
            assert int((d_st != -1).sum()) == 0
            # position-weighted checksum of the whole output (int64 wrap-around arithmetic: exact, order-independent)
            v = d_out.view(torch.int32); cs = 0
            for i0 in range(0, v.numel(), 1 << 27):
                w = v[i0:i0 + (1 << 27)].to(torch.int64)
                cs = (cs + int((w * (torch.arange(w.numel(), device=dev, dtype=torch.int64) % 1000003 + 1 + (i0 >> 27))).sum().item())) & ((1 << 63) - 1)
                del w
            if ref_sum is None: ref_sum = cs
            row[mode] = {"ms": round(ms, 2), "Mx": round(audio(a.streams) / (ms / 1e3) / 1e6, 3), "same_pcm": cs == ref_sum}
        if ne:
            def e2e():
                rc = ctx.L.ULCB_DecodeStreams(ctx.h, ne, pin_in.data_ptr(), h_offs.ctypes.data, None, a.blocks, pin_out.data_ptr(), stride,
                                              capi.FMT_PCM16, h_status.ctypes.data, None, 0)
                assert rc == 0, rc
            pin_out.zero_(); e2e()
            t0 = time.perf_counter()
            for _ in range(max(2, a.reps // 2)): e2e()
            ms = (time.perf_counter() - t0) / max(2, a.reps // 2) * 1e3
            same = bool(torch.equal(pin_out[: 64 * stride], d_out[: 64 * stride].cpu())) and bool(torch.equal(pin_out[(ne - 8) * stride:], d_out[(ne - 8) * stride: ne * stride].cpu()))
            row["e2e"] = {"streams": ne, "ms": round(ms, 1), "kx": round(audio(ne) / (ms / 1e3) / 1e3, 1), "same_pcm": same and (h_status[:ne] == -1).all().item()}
        res.append(row); print(json.dumps(row), flush=True)
