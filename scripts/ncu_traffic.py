#!/usr/bin/env python
"""Record the DRAM traffic of one launch from an `ncu --set full` report into profiles/traffic.json.

usage: python scripts/ncu_traffic.py <decode|encode> <rep.ncu-rep> <streams> <blocks>
bench.py reports it as roofline.traffic when its launch has the same shape."""
import csv, json, os, subprocess, sys
kind, rep, streams, blocks = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name); v = float(vals[i].replace(",", "")); u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
t = json.load(open(path)) if os.path.exists(path) else {}
t[kind] = {"streams": streams, "blocks": blocks, "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
           "kernel": vals[hdr.index("Kernel Name")], "source": os.path.basename(rep)}
json.dump(t, open(path, "w"), indent=1)
print(t[kind])
