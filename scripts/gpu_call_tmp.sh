set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "imfilter or conv" --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/sweep_imfilter_stagger.py > gpurun_out/r56_imfilter_regs.txt 2>&1
cat gpurun_out/r56_imfilter_regs.txt
