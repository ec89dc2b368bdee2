#!/bin/bash
# One GPU call for the Monte-Carlo math work: RNG / evolution parity tests, then the 1e8 x 256 timing in the three math modes.
set -u
TAG=${1:-r39}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "random or stochastic or monte or lcg" -p no:cacheprovider > gpurun_out/${TAG}_pytest_mc.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_mc.log
timeout 600 python scripts/time_mc.py > gpurun_out/${TAG}_mc.txt 2>&1
cat gpurun_out/${TAG}_mc.txt
