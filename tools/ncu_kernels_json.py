#!/usr/bin/env python3
"""Per-kernel summary of an .ncu-rep as JSON (last launch of each kernel name = warm instruction cache):
    tools/ncu_kernels_json.py REPORT.ncu-rep > profiles/rN_kernels_ncu_summary.json"""
import csv
import io
import json
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "launch__registers_per_thread": "registers", "launch__waves_per_multiprocessor": "waves",
        "launch__grid_size": "grid", "launch__block_size": "block",
        "sm__cycles_active.avg": "sm_cycles_active_avg", "sm__cycles_elapsed.max": "sm_cycles_elapsed_max",
        "smsp__inst_executed.sum": "warp_instructions", "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
        "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
        "sm__warps_active.avg.per_cycle_active": "active_warps_per_cycle",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_ncu_peak",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct"}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return s


out = {}
for r in data:
    name = r[col["Kernel Name"]]
    e = {"id": int(r[col["ID"]])}
    for h, k in want.items():
        if h in col:
            e[k] = num(r[col[h]])
            e[k + "_unit"] = units[col[h]]
    if "sm_cycles_active_avg" in e and "sm_cycles_elapsed_max" in e and e["sm_cycles_elapsed_max"]:
        e["sm_active_over_elapsed"] = e["sm_cycles_active_avg"] / e["sm_cycles_elapsed_max"]
    e = {k: v for k, v in e.items() if not (k.endswith("_unit") and v in ("", "%"))}
    out.setdefault(name, []).append(e)
json.dump({k: {"launches": len(v), "last": v[-1], "durations": [x.get("duration") for x in v]} for k, v in out.items()},
          sys.stdout, indent=1)
