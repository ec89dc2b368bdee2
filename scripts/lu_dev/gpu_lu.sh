#!/bin/bash
# One GPU call for the LU panel work: harness -> mldivide timing (new kernel, then the round-2 cluster kernel) -> the solve tests.
set -u
TAG=${1:-r37}
mkdir -p gpurun_out
timeout 120 scripts/lu_dev/panel_test > gpurun_out/${TAG}_lu_panel.txt 2>&1; echo "panel_test rc=$?" >> gpurun_out/${TAG}_lu_panel.txt
cat gpurun_out/${TAG}_lu_panel.txt
{
  timeout 300 python scripts/time_mldivide.py 4096
  RUNMAT_B200_LU_PANEL_V1=1 timeout 300 python scripts/time_mldivide.py 4096
  timeout 300 python scripts/time_mldivide.py 8192
  timeout 300 python scripts/time_mldivide.py 2048
} > gpurun_out/${TAG}_mldivide.txt 2>&1
cat gpurun_out/${TAG}_mldivide.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "mldivide or linsolve or mrdivide" -p no:cacheprovider > gpurun_out/${TAG}_pytest_lu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_lu.log
