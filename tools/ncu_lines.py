#!/usr/bin/env python
"""Per CUDA source line: warp-instructions executed, average active lanes and stall samples, from an
.ncu-rep captured with --import-source on (kernel compiled with -lineinfo).
Usage: python tools/ncu_lines.py report.ncu-rep [launch-index] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, lines = "?", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    off = len(r) - len(hdr)  # unescaped quotes/commas in the source text shift the numeric columns
    try:
        ie = int(r[off + hdr["Instructions Executed"]] or 0)
        te = int(r[off + hdr["Thread Instructions Executed"]] or 0)
        sm = int(r[off + hdr["# Samples"]] or 0)
    except ValueError:
        continue
    lines.append((ie, te, sm, fname, int(r[0]), ",".join(r[1:2 + off]).strip()))
tot_i = sum(l[0] for l in lines) or 1
tot_s = sum(l[2] for l in lines) or 1
print(f"total warp-instructions {tot_i/1e6:.1f} M, stall samples {tot_s}")
print("| warp-inst (M) | share | lanes | samples | where | source |\n|---|---|---|---|---|---|")
for ie, te, sm, fn, ln, src in sorted(lines, key=lambda l: -l[0])[:top]:
    print(f"| {ie/1e6:.1f} | {100*ie/tot_i:.1f}% | {te/max(ie,1):.1f} | {100*sm/tot_s:.1f}% | {fn}:{ln} | `{src[:90]}` |")
