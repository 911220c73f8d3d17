#!/bin/bash
# Runs on the GPU box (gpurun, ONE GPU): launch list of one bench step, per-launch DRAM traffic of the GEMM
# kernels, and `ncu --set full` captures of the hot kernels.  Outputs go to gpurun_out/; tools/write_profiles.sh
# turns them into the text files committed under profiles/.
# usage: bash tools/refresh_profiles.sh <tag>      (e.g. r02)
set -u
tag=${1:-r02}
out=gpurun_out
# the bench's own command (32-view job, 8 views per decoder pass); eager launches: same kernels, simple launch order
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --no-cuda-graphs"
timeout 300 $B > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err || { echo "plain bench failed"; exit 1; }
L=$(python -c "import json; d=json.load(open('$out/${tag}_bench_plain.json')); print(d['gpu_launches'] // d['steps'])")
echo "launches per step: $L"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3 * L)) -c $L --csv \
    --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k "regex:gemm_tc|conv_halo" --launch-skip 500 -c 400 --csv --log-file $out/${tag}_gemm_traffic.csv $B > $out/${tag}_ncu_traffic.log 2>&1
# one scene + one 8-view decoder pass (the bench's chunk) for the full captures
P="python tools/profile_step.py --views 8"
timeout 120 $P > $out/${tag}_profile_step.log 2>&1 || { echo "profile_step failed"; exit 1; }
# GEMM launch order inside a step (Large): 51 scene-stage GEMMs, then the ray-token GEMM, then per decoder layer
# {wq (PROJ16), wout (RESID), s.wqkv (PROJ16 + transposed V tiles), s.wo (RESID), w13 (SwiGLU), w2 (RESID)}:
# skipping 70 lands on decoder layer 3 and 6 launches cover one instance of each epilogue kind
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 70 -c 6 -o $out/${tag}_gemm_layer $P > $out/${tag}_ncu1.log 2>&1
# attention launch order: 12 x attn2 (encoder), then per decoder layer attn3 (cross) and attn_swin:
# skipping 11 captures the last encoder layer's attn2, the first cross-attention and the first swin layer
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attn2_tc|attn3_tc|attn_swin" --launch-skip 11 -c 3 -o $out/${tag}_attn $P > $out/${tag}_ncu2.log 2>&1
ls -la $out | grep ${tag}_
