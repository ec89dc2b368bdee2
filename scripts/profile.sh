#!/bin/bash
# Runs on the GPU box (under gpurun). Captures: (1) launch list with device times for one bench command,
# (2) ncu --set full of the dominant kernels. Outputs land in gpurun_out/ and are summarised into profiles/ on the CPU box.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
BENCH="python bench.py --steps 4 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_fused_ew -s 4 -c 2 -f -o gpurun_out/${TAG}_fused_ew python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rm_fused_red -s 4 -c 2 -f -o gpurun_out/${TAG}_fused_red python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ozaki_gemm|slice_kernel|dgemm_dmma" -c 5 -f -o gpurun_out/${TAG}_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"evolve_kernel|lu_panel_push" -c 3 -f -o gpurun_out/${TAG}_mc_lu python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"moments|normalize|imfilter" -c 8 -f -o gpurun_out/${TAG}_image python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"imfilter" -c 3 -f -o gpurun_out/${TAG}_imfilter python scripts/time_imfilter.py > gpurun_out/${TAG}_imfilter_ncu.log 2>&1
ls -la gpurun_out
