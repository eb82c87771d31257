#!/usr/bin/env python
"""Aggregate an ncu report's source page by CUDA source line: python scripts/ncu_lines.py <rep> [top]"""
import csv, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; out = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-": continue     # line-level rows have "-" in the Address column
    g = lambda n: int(r[hdr.index(n)] or 0) if r[hdr.index(n)].lstrip("-").isdigit() else 0
    out.append((g("# Samples"), g("Instructions Executed"), cur, r[0], r[1].strip(), {k: g(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_lg", "stall_math", "stall_not_selected", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal")}))
ts = sum(o[0] for o in out); ti = sum(o[1] for o in out)
print(f"total samples {ts}, total warp-instructions {ti}")
by_file = {}
for o in out: by_file.setdefault(o[2], [0, 0]); by_file[o[2]][0] += o[0]; by_file[o[2]][1] += o[1]
for f, (s, i) in by_file.items(): print(f"  {f}: {100*s/ts:.1f}% samples, {100*i/ti:.1f}% instructions")
out.sort(key=lambda o: -o[1] if os.environ.get("BY_INSTR") else -o[0])
for s, i, f, ln, src, st in out[:top]:
    big = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*s/ts:5.1f}%smp {100*i/ti:5.1f}%ins {f[:14]:>14}:{ln:>4} {src[:80]:80} | " + " ".join(f"{k.replace('stall_','')}={v}" for k, v in big if v))
