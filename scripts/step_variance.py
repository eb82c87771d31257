import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch, bench
from ulc_codec_b200 import capi
blocks = bench.load_pool(); dev = torch.device("cuda", 0)
distinct, index = bench.make_streams(blocks, 256, 0)
b = bench.DecodeBatch(torch, dev, distinct, index, 4096)
d_out = torch.empty(4096 * b.stride, dtype=torch.int16, device=dev)
with capi.Context(44100, 2, 2048, max_blocks=bench.BLOCKS_PER_STREAM) as ctx:
    ext = torch.cuda.ExternalStream(ctx.cuda_stream, device=dev)
    ctx.set_kernel_timing(True)
    for mode in ("index", "walk", "index", "walk"):
        ts = []
        for i in range(6):
            kw = dict(d_block_bytes=b.d_index.data_ptr(), block_bytes_stride=bench.BLOCKS_PER_STREAM) if mode == "index" else {}
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            ctx.decode_streams_device(4096, b.d_bytes.data_ptr(), b.d_offs.data_ptr(), bench.BLOCKS_PER_STREAM, d_out.data_ptr(), b.stride, "pcm16", d_status=b.d_status.data_ptr(), **kw)
            e1.record(ext); ctx.sync(); torch.cuda.synchronize()
            kt = ctx.kernel_times()
            ts.append((round(e0.elapsed_time(e1), 1), round(kt["bp_parse"][0], 1), round(kt["bp_transform"][0], 1), round(kt["bp_walk"][0] + kt["bp_index"][0], 1)))
        print(mode, ts)
