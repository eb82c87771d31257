#!/usr/bin/env python
"""Instruction / sample share per source region: python scripts/ncu_regions.py <rep> file:lo-hi=name ..."""
import csv, subprocess, sys
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    loc, name = a.split("="); f, rng = loc.split(":"); lo, hi = rng.split("-"); regions.append((f, int(lo), int(hi), name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; acc = {}; ts = ti = 0
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue
    try: ln = int(r[0]); s = int(r[hdr.index("# Samples")] or 0); i = int(r[hdr.index("Instructions Executed")] or 0)
    except ValueError: continue
    ts += s; ti += i
    name = f"{cur}:other"
    for f, lo, hi, nm in regions:
        if cur.startswith(f) and lo <= ln <= hi: name = nm; break
    a = acc.setdefault(name, [0, 0]); a[0] += s; a[1] += i
for k, (s, i) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34} {100*i/ti:5.1f}% instr ({i/1e6:9.1f}M)  {100*s/ts:5.1f}% samples")
print("total warp-instr", ti)
