#!/usr/bin/env python
"""Aggregate warp-stall samples of an ncu report by stall reason and by CUDA source line.
usage: python tools/ncu_stalls.py X.ncu-rep [top_lines]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
by_reason, by_line, ninstr = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0.0
hdr, ci, stalls, fname = None, None, None, ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {}
        for i, h in enumerate(hdr):
            ci.setdefault(h, i)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[2] == "-":
        continue
    try:
        v = float(r[ci["# Samples"]])
    except ValueError:
        continue
    tot += v
    key = (fname, r[0], r[1].strip()[:100])
    by_line[key] += v
    ninstr[key] += 1
    for s in stalls:
        try:
            by_reason[s[6:]] += float(r[ci[s]] or 0)
        except ValueError:
            pass
print("total samples", tot)
print("by reason: " + "  ".join(f"{k}:{100 * v / max(tot, 1):.1f}%" for k, v in by_reason.most_common(10)))
for (f, ln, src), v in by_line.most_common(top):
    print(f"{v:7.0f} {100 * v / tot:5.1f}%  {f}:{ln:>4s} [{ninstr[(f, ln, src)]:3d} sass]  {src}")
