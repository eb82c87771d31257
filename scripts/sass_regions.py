#!/usr/bin/env python
"""Static SASS size per source line of one kernel: python scripts/sass_regions.py <obj.o> <kernel substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
td = tempfile.mkdtemp(); subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, capture_output=True)
cub = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur = None; fn = None; cnt = collections.Counter(); srcf = None
for l in out.splitlines():
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m: srcf, cur = m.group(1), int(m.group(2)); continue
    m = re.match(r'\.text\.(\S+):', l)
    if m: fn = m.group(1); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and fn and pat in fn: cnt[(os.path.basename(srcf or "?"), cur)] += 1
tot = sum(cnt.values()); print("total SASS instructions:", tot)
src = {}
for (f, ln), v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        try: src[f] = open(os.path.join(os.path.dirname(obj), "..", "csrc", f)).read().splitlines()
        except OSError: src[f] = []
    text = src[f][ln - 1].strip()[:100] if ln and ln <= len(src[f]) else ""
    print(f"{v:6d} {100*v/tot:5.1f}% {f}:{ln} {text}")
