#!/usr/bin/env python
"""Summarise every kernel of an `ncu --set full --import-source on` report: key metrics, stall mix, hottest source lines.
usage: python scripts/ncu_report.py <rep.ncu-rep> [top_lines]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
print(f"# {rep}")
for r in rows[2:]:
    print("## kernel")
    for wname in want:
        if wname in hdr: print(f"{wname} = {r[hdr.index(wname)]} {rows[1][hdr.index(wname)]}")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h:
            try: st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    print("stalls per issue: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True) if v > 0.05))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
h = None; data = []
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Line No": h = r; continue
    if h is None or len(r) < len(h) or r[2] != "-": continue
    try: data.append((int(r[0]), int(r[h.index("Instructions Executed")] or 0), int(r[h.index("# Samples")] or 0), r[1].strip()))
    except ValueError: pass
ti = sum(x[1] for x in data); ts = sum(x[2] for x in data)
print(f"## hottest source lines (all kernels of the report; {ti} warp-instructions, {ts} samples)")
for ln, i, s, src in sorted(data, key=lambda x: -(x[2] if "--by-samples" in sys.argv else x[1]))[:top]:
    print(f"{100 * i / max(ti, 1):5.1f}%ins {100 * s / max(ts, 1):5.1f}%smp {ln:4d} {src[:120]}")
