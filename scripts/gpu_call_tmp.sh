set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "imfilter or conv" --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python scripts/time_imfilter.py > gpurun_out/r59_imfilter.txt 2>&1
grep "tma-staged packed rbw8\|flags" gpurun_out/r59_imfilter.txt
