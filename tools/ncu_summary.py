#!/usr/bin/env python3
"""Summarise an .ncu-rep: key launch metrics + instructions by source line.
usage: tools/ncu_summary.py REPORT.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu",
        "smsp__warps_eligible.avg.per_cycle_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes_pipe_lsu_mem_local",
        "smsp__inst_executed_op_local", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
for i, h in enumerate(hdr):
    if any(h == w or h.startswith(w + ".") and h == w for w in want) or h in want:
        print("%-72s %s" % (h, [r[i] for r in rows[1:]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = func = None
agg, seen = [], set()
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1][:50]
        continue
    if r[0].isdigit():
        try:
            key = (cur, int(r[0]), func)
            if key in seen:
                continue
            seen.add(key)
            agg.append((int(r[7]), int(r[8]), int(r[6]), cur, int(r[0]), r[1].strip()[:88]))
        except Exception:
            pass
tot = sum(a[0] for a in agg) or 1
print("total warp instructions (first kernel instance): %d" % tot)
for a in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% inst=%9d thr/inst=%5.1f samples=%5d %s:%d  %s" % (100 * a[0] / tot, a[0], a[1] / max(a[0], 1), a[2], a[3], a[4], a[5]))
