#!/usr/bin/env python
"""Sum the DRAM traffic of one decode step (every kernel of the block-parallel pipeline, every wave) from an ncu CSV launch list
captured with --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, and record it in profiles/traffic.json
keyed by launch shape and build (bench.py reports it as roofline.traffic when both match).

usage: python scripts/ncu_traffic_bp.py <launches.csv> <streams> <blocks> <steps-in-capture> [decode|encode]
(encode: the detector pre-pass + the encode kernel of one ULCB_EncodeStreamsDevice call)"""
import csv, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
path, streams, blocks, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
family = sys.argv[5] if len(sys.argv) > 5 else "decode"
wanted = ("bp_",) if family == "decode" else ("ulcb_encode_kernel", "ulcb_detect_kernel")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}
per = {}
for r in rows[1:]:
    if not any(w in r[ik] for w in wanted):
        continue
    k = r[ik].split("::")[-1].split("(")[0]
    d = per.setdefault(k, {"dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "ns": 0.0, "launches": 0})
    v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1)
    if r[im] == "dram__bytes_read.sum": d["dram_bytes_read"] += v
    elif r[im] == "dram__bytes_write.sum": d["dram_bytes_write"] += v
    elif r[im] == "gpu__time_duration.sum": d["ns"] += v; d["launches"] += 1
for d in per.values():
    for k in d: d[k] /= steps
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
t = json.load(open(out)) if os.path.exists(out) else {}
t[family] = {"streams": streams, "blocks": blocks, "build": bench.build_id(), "per_step": True,
               "dram_bytes_read": sum(d["dram_bytes_read"] for d in per.values()), "dram_bytes_write": sum(d["dram_bytes_write"] for d in per.values()),
               "kernels": per, "source": os.path.basename(path)}
json.dump(t, open(out, "w"), indent=1)
print(json.dumps(t[family], indent=1))
