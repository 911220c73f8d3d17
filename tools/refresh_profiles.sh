#!/bin/bash
# Runs on the GPU box (gpurun): launch list, per-launch DRAM traffic of the GEMM kernel and
# `ncu --set full` captures of the hot kernels of one bench step.  Outputs go to gpurun_out/;
# tools/ncu_summary.py / ncu_stalls.py turn them into the text files committed under profiles/.
# usage: bash tools/refresh_profiles.sh <tag>      (e.g. r01b)
set -u
tag=${1:-r01}
out=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cuda-graphs"   # eager launches: same kernels, simple launch order
$B > $out/${tag}_bench_plain.json 2> $out/${tag}_bench_plain.err || { echo "plain bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 215 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:gemm_tc|conv_halo" \
    --launch-skip 155 -c 155 --csv --log-file $out/${tag}_gemm_traffic.csv $B > $out/${tag}_ncu_traffic.log 2>&1
P="python tools/profile_step.py"
# GEMM launch order inside a step (Large): 51 scene-stage GEMMs, then the ray-token GEMM, then per decoder
# layer {wq (PROJ16), wout (RESID), s.wqkv (PROJ16 + transposed V tiles), s.wo (RESID), w13 (SwiGLU), w2 (RESID)}:
# skipping 70 lands on decoder layer 3 and 6 launches cover one instance of each epilogue kind
ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 70 -c 6 -o $out/${tag}_gemm_layer $P > $out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn3_tc --launch-skip 2 -c 1 -o $out/${tag}_attn_cross $P > $out/${tag}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn2_tc --launch-skip 2 -c 1 -o $out/${tag}_attn_enc $P > $out/${tag}_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_swin --launch-skip 2 -c 1 -o $out/${tag}_attn_swin $P > $out/${tag}_ncu5.log 2>&1
ls -la $out | grep ${tag}_
