#!/usr/bin/env python
"""Attributes every SASS instruction of a profiled kernel to ONE source line (the innermost frame that lies in
the files given in PREFER order) and prints warp-instructions, average active lanes and stall samples per line
and per named line range.  Needs --import-source on and -lineinfo.
Usage: python tools/ncu_regions.py report.ncu-rep [launch-index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
PREFER = ["trace_sm.cuh", "traverse.cuh", "shade.cuh", "kernels.cu"]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, cur = "?", None, None
inst = {}  # address -> (rank, file, line, ie, te, samples, sass)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        continue
    if hdr is None:
        continue
    if r[0].strip().isdigit():
        cur = int(r[0])
        continue
    if len(r) < len(hdr) or not r[2].startswith("0x"):
        continue
    addr = r[2]
    try:
        ie, te, sm = int(r[hdr["Instructions Executed"]] or 0), int(r[hdr["Thread Instructions Executed"]] or 0), int(r[hdr["# Samples"]] or 0)
    except ValueError:
        continue
    rank = PREFER.index(fname) if fname in PREFER else len(PREFER)
    if addr not in inst or rank < inst[addr][0]:
        inst[addr] = (rank, fname, cur, ie, te, sm, r[3].strip())
tot_i = sum(v[3] for v in inst.values()) or 1
tot_t = sum(v[4] for v in inst.values())
tot_s = sum(v[5] for v in inst.values()) or 1
print(f"{len(inst)} SASS instructions, {tot_i/1e6:.1f} M warp-inst, avg lanes {tot_t/tot_i:.2f}, {tot_s} samples")
by_line = collections.defaultdict(lambda: [0, 0, 0])
for rank, fn, ln, ie, te, sm, _ in inst.values():
    a = by_line[(fn, ln)]
    a[0] += ie
    a[1] += te
    a[2] += sm
print("\n| where | warp-inst (M) | share | lanes | samples |\n|---|---|---|---|---|")
for (fn, ln), (ie, te, sm) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:50]:
    print(f"| {fn}:{ln} | {ie/1e6:.1f} | {100*ie/tot_i:.1f}% | {te/max(ie,1):.1f} | {100*sm/tot_s:.1f}% |")
by_file = collections.defaultdict(lambda: [0, 0, 0])
for rank, fn, ln, ie, te, sm, _ in inst.values():
    a = by_file[fn]
    a[0] += ie
    a[1] += te
    a[2] += sm
print("\n| file | warp-inst (M) | share | lanes | samples |\n|---|---|---|---|---|")
for fn, (ie, te, sm) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"| {fn} | {ie/1e6:.1f} | {100*ie/tot_i:.1f}% | {te/max(ie,1):.1f} | {100*sm/tot_s:.1f}% |")
