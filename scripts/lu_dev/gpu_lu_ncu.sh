#!/bin/bash
# harness -> (if it passes) mldivide timing + solve / pattern tests -> ncu of the m = 4096 and m = 64 panel launches with dense warp sampling
set -u
TAG=${1:-r42}
mkdir -p gpurun_out
timeout 120 scripts/lu_dev/panel_test > gpurun_out/${TAG}_lu_panel.txt 2>&1; echo "panel_test rc=$?" >> gpurun_out/${TAG}_lu_panel.txt
cut -c1-100,230-300 gpurun_out/${TAG}_lu_panel.txt
if grep -q "ALL PASS" gpurun_out/${TAG}_lu_panel.txt; then
  timeout 300 python scripts/time_mldivide.py 4096 > gpurun_out/${TAG}_mldivide.txt 2>&1; cat gpurun_out/${TAG}_mldivide.txt
  timeout 600 python -m pytest tests -m gpu -q -x -k "mldivide or linsolve or mrdivide or syrk or pattern" -p no:cacheprovider > gpurun_out/${TAG}_pytest_lu.log 2>&1
  tail -5 gpurun_out/${TAG}_pytest_lu.log
fi
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:lu_panel_push -s 110 -c 1 -f -o gpurun_out/${TAG}_lu_panel scripts/lu_dev/panel_test > gpurun_out/${TAG}_lu_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:lu_panel_push -s 2 -c 1 -f -o gpurun_out/${TAG}_lu_panel_m64 scripts/lu_dev/panel_test >> gpurun_out/${TAG}_lu_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_lu_ncu.log
