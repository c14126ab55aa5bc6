#!/bin/bash
# round-2 profile capture (final kernels): launch list of the bench command, full capture of the two iteration kernels
# (config 2), full capture of the config-4 iteration kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PDLP_BENCH_CFG4=0
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 380 -c 2 -o gpurun_out/r2_iter -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
echo "full cfg2 rc=$?"
export PDLP_TIMING_ONLY=1
python scripts/dev_time.py cfg4 > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'stream_kernel<(Primal|Dual)' -s 60 -c 2 -o gpurun_out/r2_cfg4 -f python scripts/dev_time.py cfg4 > gpurun_out/ncu4.log 2>&1
echo "full cfg4 rc=$?"
tail -n 3 gpurun_out/ncu_full.log gpurun_out/ncu4.log
