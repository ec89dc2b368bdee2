set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --timeout 600 -p no:cacheprovider > gpurun_out/r50_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r50_pytest.log
tail -15 gpurun_out/r50_pytest.log
timeout 600 python scripts/sweep_reduction.py > gpurun_out/r50_red_sweep.txt 2>&1
cat gpurun_out/r50_red_sweep.txt
