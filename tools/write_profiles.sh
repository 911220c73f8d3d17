#!/bin/bash
# Turns the raw outputs of tools/refresh_profiles.sh <tag> (gpurun_out/) into the text files under profiles/.
# usage: bash tools/write_profiles.sh <tag>
set -u
T=${1:-r02}
cp gpurun_out/${T}_launches.csv profiles/${T}_launches.csv
python tools/ncu_summary.py launches gpurun_out/${T}_launches.csv > profiles/${T}_launches_summary.txt
cp gpurun_out/${T}_profile_step.log profiles/${T}_profile_step.txt
{ echo "# ncu --set full, one decoder layer (layer 3) of one scene + 8-view pass: wq (PROJ16), wout (RESID), s.wqkv (PROJ16 + transposed V tiles), s.wo (RESID + row map), w13 (SwiGLU), w2 (RESID, K=4096)"; python tools/ncu_summary.py report gpurun_out/${T}_gemm_layer.ncu-rep; } > profiles/${T}_gemm_layer_summary.txt
{ echo "# ncu --set full: encoder self-attention (8 heads, 4112 tokens; attn2_tc_kernel, two query tiles per CTA), decoder cross-attention (B=8 views, 8 heads, Nq=4096, Nk=4112), swin window attention (32768 tokens, 8 heads)"; python tools/ncu_summary.py report gpurun_out/${T}_attn.ncu-rep; for i in 1 2 3; do python tools/ncu_stalls.py gpurun_out/${T}_attn.ncu-rep $i 2>/dev/null | head -2; done; } > profiles/${T}_attn_summary.txt
python - "$T" <<'PY'
import csv, collections, json, sys
T = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/{T}_gemm_traffic.csv") if l.startswith('"'))]
hdr = rows[0]
mi, vi, ui, ii = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = collections.defaultdict(dict)
for r in rows[1:]:
    v, u = float(r[vi].replace(",", "")), r[ui]
    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u] if r[mi].startswith("dram") else {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    per[r[ii]][r[mi]] = v
n = len(per)
rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values())
wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values())
t = sum(p.get("gpu__time_duration.sum", 0) for p in per.values())
d = {"kernel": "gemm_tc_kernel / conv_halo_kernel (all template instances)", "launches": n, "dram_bytes_read_per_launch": rd / n,
     "dram_bytes_write_per_launch": wr / n, "dram_bytes_per_launch": (rd + wr) / n, "avg_duration_us_under_ncu": t / n,
     "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc|conv_halo "
               f"--launch-skip 500 -c 400 python bench.py --steps 2 --warmup 3 --no-cuda-graphs "
               f"(tools/refresh_profiles.sh {T}): 400 consecutive GEMM / conv launches of the 32-view job"}
json.dump(d, open(f"profiles/{T}_gemm_traffic.json", "w"), indent=1)
print(n, "GEMM launches,", round(d["dram_bytes_per_launch"] / 1e6, 1), "MB per launch,", round(d["avg_duration_us_under_ncu"], 1), "us avg")
PY
