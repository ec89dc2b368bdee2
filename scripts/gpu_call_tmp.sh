set -u
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r61_bench.json 2> gpurun_out/r61_bench.err
echo "bench rc=$? lines=$(wc -l < gpurun_out/r61_bench.json)"; head -c 400 gpurun_out/r61_bench.json; echo
timeout 60 ncu --set full --clock-control none --import-source on -k regex:imfilter -s 2 -c 1 -f -o gpurun_out/r61_imfilter5 python scripts/ncu_imfilter5.py > gpurun_out/r61_imfilter5_ncu.log 2>&1
echo "ncu imfilter rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r61_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r61_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:rm_fused_ew -s 4 -c 2 -f -o gpurun_out/r61_fused_ew python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
timeout 60 ncu --set full --clock-control none --import-source on -k regex:rm_fused_red -s 4 -c 2 -f -o gpurun_out/r61_fused_red python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
ls gpurun_out | grep r61
