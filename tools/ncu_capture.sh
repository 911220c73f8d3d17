set -x
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-roofline"
$CMD > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1131 -c 377 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 606 -c 6 -o gpurun_out/r01_gemm_full $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 120 -c 3 -o gpurun_out/r01_attn_full $CMD > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/
