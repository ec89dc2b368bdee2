set -u
mkdir -p gpurun_out
SKIP_PROFILE=1 bash scripts/gpu_round.sh r58 > gpurun_out/r58_round.log 2>&1
tail -4 gpurun_out/r58_pytest.log
timeout 300 python scripts/time_imfilter.py > gpurun_out/r58_imfilter.txt 2>&1
grep "tma-staged packed rbw8\|flags" gpurun_out/r58_imfilter.txt
RUNMAT_B200_NO_PDL=1 timeout 300 python scripts/time_imfilter.py 2>&1 | grep "tma-staged packed rbw8" | sed 's/^/NO_PDL /' | tee -a gpurun_out/r58_imfilter.txt
timeout 300 python scripts/time_image_normalize.py > gpurun_out/r58_image_normalize.txt 2>&1
cat gpurun_out/r58_image_normalize.txt
RUNMAT_B200_NO_PDL=1 timeout 300 python scripts/time_image_normalize.py 2>&1 | head -2 | sed 's/^/NO_PDL /' | tee -a gpurun_out/r58_image_normalize.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r58_bench.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['reduction_kernel']['us_per_launch'], d['roofline']['pipelined'])
print('e2e', d['e2e']['ms_per_step'])
for k,v in d['extra'].items(): print(k, v.get('ms') or v.get('ms_per_batch'), v['roofline']['frac'])
PY
