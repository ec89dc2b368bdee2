set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r60_bench_n2.json 2> gpurun_out/r60_bench_n2.err
echo "bench n2 rc=$?"; tail -5 gpurun_out/r60_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r60_bench_n2.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['run']['exchange'], d['run']['exchange_verified'])
print('e2e', d['e2e']['ms_per_step'])
for k,v in d['extra'].items(): print(k, v.get('ms') or v.get('ms_per_batch'), v.get('exchange'), v.get('price') or v.get('mse'))
PY
