#!/usr/bin/env python
"""Summarise ncu outputs into text files for profiles/:
   launches:  python tools/ncu_summary.py launches gpurun_out/r01_launches.csv
   report:    python tools/ncu_summary.py report gpurun_out/r01_gemm_full.ncu-rep"""
import collections
import csv
import subprocess
import sys

mode, path = sys.argv[1], sys.argv[2]
if mode == "launches":
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v  # -> us
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    print(f"# {path}: {len(rows) - 1} launches, {tot / 1e3:.3f} ms summed kernel time (cold-cache, serialised)")
    for k, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{us / 1e3:9.3f} ms {100 * us / tot:5.1f}%  n={n:4d}  avg {us / n:8.1f} us  {k}")
else:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
            "sm__cycles_elapsed.max"]
    units = rows[1]
    for r in rows[2:]:
        print("-" * 100)
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:70s} {r[i][:60]:>20s} {units[i]}")
