#!/bin/bash
# pattern / solve tests, then the warm launch list of ONE mldivide (n = 4096, 64 rhs) with device times per kernel and stream
set -u
TAG=${1:-r43}
mkdir -p gpurun_out
timeout 120 scripts/lu_dev/panel_test > gpurun_out/${TAG}_lu_panel.txt 2>&1; echo "panel_test rc=$?" >> gpurun_out/${TAG}_lu_panel.txt
grep -o "PASS.*info=.\|FAIL.*info=.\|time median.*\|ALL PASS\|FAILED.*" gpurun_out/${TAG}_lu_panel.txt | paste - - | head -20
timeout 300 python scripts/time_mldivide.py 4096 > gpurun_out/${TAG}_mldivide.txt 2>&1; cat gpurun_out/${TAG}_mldivide.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "mldivide or linsolve or mrdivide or syrk or pattern" -p no:cacheprovider > gpurun_out/${TAG}_pytest_lu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_lu.log
cat > /tmp/one_solve.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from runmat_b200 import B200Provider
n = 4096
p = B200Provider(0)
rng = np.random.default_rng(0)
A = rng.uniform(-1, 1, (n, n)) + np.eye(n) * 4.0
B = rng.uniform(-1, 1, (n, 64))
hM, hR = p.upload(A), p.upload(B)
for _ in range(3):
    p.free(p.mldivide(hM, hR)); p.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/${TAG}_mldivide_launches.csv python /tmp/one_solve.py > gpurun_out/${TAG}_mldivide_launches.log 2>&1
wc -l gpurun_out/${TAG}_mldivide_launches.csv
