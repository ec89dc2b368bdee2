#!/bin/bash
# One 8-GPU call: the N = 4 and N = 8 bench lines (peer-memory exchange) under torchrun.
#   gpurun --gpus 8 --timeout 1200 -- 'bash scripts/gpu_scale.sh r20'
set -u
TAG=${1:-r20}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
for N in ${NLIST:-4 8}; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  echo "bench n$N rc=$?"; tail -3 gpurun_out/${TAG}_bench_n$N.err; grep '^{' gpurun_out/${TAG}_bench_n$N.json | head -c 400; echo
done
