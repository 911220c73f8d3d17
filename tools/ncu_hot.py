#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data, tot = [], 0.0
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr) or r[0] == "Address":
        break
    try:
        v = float(r[ci["# Samples"]])
    except ValueError:
        continue
    tot += v
    why = sorted(((float(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    data.append((v, idx, r[ci["Source"]].strip()[:90], why, r[ci["Instructions Executed"]]))
print("total samples", tot)
for v, idx, s, why, n in sorted(data, reverse=True)[:top]:
    w = " ".join(f"{b}:{a:.0f}" for a, b in why if a > 0)
    print(f"{v:8.0f} {100 * v / tot:5.1f}%  #{idx:5d} x{n:>8}  {s:90s} {w}")
