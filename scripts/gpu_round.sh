#!/bin/bash
# One GPU call: parity tests -> bench (both arms) -> launch list -> ncu captures. Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh r05'
set -u
TAG=${1:-r05}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "${SKIP_PROFILE:-0}" != "1" ]; then
  timeout 900 bash scripts/profile.sh ${TAG} > gpurun_out/${TAG}_profile.log 2>&1
fi
ls -la gpurun_out | tail -20
