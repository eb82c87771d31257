#!/usr/bin/env python
"""Summarise one `ncu --set full` report into a text file for profiles/.

usage: python scripts/ncu_summary.py <rep.ncu-rep> [--lines N] > profiles/rNN_x_ncu_summary.txt
Prints the raw metrics the roofline uses (duration, DRAM bytes, pipe utilisation, occupancy limits,
stall mix) and the hottest source lines by sampled stalls.
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 30

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
print(f"# {rep}")
kn = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
if kn is not None:
    print(f"# kernel: {vals[kn]}")
for m in WANT:
    if m in hdr:
        i = hdr.index(m)
        print(f"{m} = {vals[i]} {units[i]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur = None; h = None; lines = []; ts = ti = 0
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": h = r; continue
    if h is None or len(r) < len(h) or r[2] != "-": continue
    try:
        ln = int(r[0]); s = int(r[h.index("# Samples")] or 0); i = int(r[h.index("Instructions Executed")] or 0)
    except ValueError:
        continue
    ts += s; ti += i
    lines.append((s, i, cur, ln, r[1].strip()[:90]))
print(f"\n## hottest source lines (of {ts} samples, {ti} warp-instructions)")
for s, i, f, ln, txt in sorted(lines, key=lambda x: -x[0])[:nlines]:
    print(f"{100*s/max(ts,1):5.1f}%smp {100*i/max(ti,1):5.1f}%ins {f}:{ln}  {txt}")
