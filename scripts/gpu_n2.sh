#!/bin/bash
# One 2-GPU call: new parity tests on GPU 0, the single-GPU bench, then the N=2 bench (peer-memory exchange) under torchrun.
#   gpurun --gpus 2 --timeout 1500 -- 'bash scripts/gpu_n2.sh r07'
set -u
TAG=${1:-r07}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout 600 -p no:cacheprovider -k "${KEXPR:-reduc or interleaved or strided or image or normalize or comm or mean}" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "bench n$N rc=$?"; tail -5 gpurun_out/${TAG}_bench_n$N.err; head -c 600 gpurun_out/${TAG}_bench_n$N.json
