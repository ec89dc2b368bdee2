#!/bin/bash
# One GPU call for the LU panel work: harness -> mldivide timing (new kernel, then the round-2 cluster kernel) -> the solve tests
# (-> optional ncu capture of the Monte-Carlo kernel, bundled to save a box acquisition).
set -u
TAG=${1:-r37}
mkdir -p gpurun_out
timeout 120 scripts/lu_dev/panel_test > gpurun_out/${TAG}_lu_panel.txt 2>&1; echo "panel_test rc=$?" >> gpurun_out/${TAG}_lu_panel.txt
cat gpurun_out/${TAG}_lu_panel.txt
if grep -q "ALL PASS" gpurun_out/${TAG}_lu_panel.txt; then
{
  timeout 300 python scripts/time_mldivide.py 4096
  timeout 300 python scripts/time_mldivide.py 8192
  timeout 300 python scripts/time_mldivide.py 2048
} > gpurun_out/${TAG}_mldivide.txt 2>&1
cat gpurun_out/${TAG}_mldivide.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "mldivide or linsolve or mrdivide" -p no:cacheprovider > gpurun_out/${TAG}_pytest_lu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_lu.log
fi
if [ "${NCU_MC:-0}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:evolve_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_mc python scripts/time_mc.py > gpurun_out/${TAG}_mc_ncu.log 2>&1
  tail -3 gpurun_out/${TAG}_mc_ncu.log
fi
