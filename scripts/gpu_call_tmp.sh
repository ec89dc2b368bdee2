set -u
mkdir -p gpurun_out
timeout 300 python scripts/sweep_reduction.py > gpurun_out/r57_red_sweep.txt 2>&1
cat gpurun_out/r57_red_sweep.txt
bash scripts/gpu_round.sh r57
