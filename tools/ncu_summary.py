#!/usr/bin/env python
"""Condenses one .ncu-rep (ncu --set full --import-source on) into a short text summary:
key launch / throughput / stall metrics, an opcode-class table with lane occupancy, and the hottest SASS lines.
Usage: python tools/ncu_summary.py report.ncu-rep [launch-index] > profiles/<name>.md"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"   # which launch of the report
SEL = ["--launch-skip", skip, "--launch-count", "1"]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + SEL, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
print(f"# ncu summary of {rep}\n")
m = dict(zip(hdr, zip(units, vals)))
print("kernel:", m.get("Kernel Name", ("", "?"))[1][:160], "\n")
print("| metric | value | unit |\n|---|---|---|")
for k in KEYS:
    if k in m:
        print(f"| {k} | {m[k][1]} | {m[k][0]} |")
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        v = float(m[k][1] or 0)
        if v >= 0.05:
            print(f"| stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} | {v:.3f} | warps/issue |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + SEL, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) >= len(h) and r[0].startswith("0x")]
cls, thr = collections.Counter(), collections.Counter()
for r in data:
    ie = int(r[ix["Instructions Executed"]] or 0)
    te = int(r[ix["Thread Instructions Executed"]] or 0)
    mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = mm.group(2) if mm else "?"
    cls[op] += ie
    thr[op] += te
tot = sum(cls.values())
print(f"\nSASS lines {len(data)}, warp-instructions executed {tot}\n")
print("| opcode | warp-inst (M) | share | avg active lanes |\n|---|---|---|---|")
for op, c in cls.most_common(16):
    print(f"| {op} | {c/1e6:.2f} | {100*c/tot:.1f}% | {thr[op]/max(c,1):.1f} |")
print("\nhottest SASS by stall samples:\n")
samp = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]
tots = sum(int(r[ix["# Samples"]] or 0) for r in data)
for r in samp:
    print(f"    {int(r[ix['# Samples']] or 0):6d} ({100*int(r[ix['# Samples']] or 0)/max(tots,1):4.1f}%) lanes {float(r[ix['Avg. Threads Executed']] or 0):4.1f}  {r[ix['Source']].strip()[:100]}")
