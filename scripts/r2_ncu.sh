#!/bin/bash
# developer run: ncu --set full on the new kernels (config 2), panel layout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PDLP_TIMING_ONLY=1
PDLP_LAYOUT=${1:-panel} timeout 300 python scripts/dev_time.py cfg2 > gpurun_out/ncu_plain.log 2>&1 &&
PDLP_LAYOUT=${1:-panel} timeout 900 ncu --set full --clock-control none --import-source on -k regex:'run_kernel|panel_kernel|combine_kernel' -s 60 -c 3 \
    -o gpurun_out/r2_prof_${1:-panel} -f python scripts/dev_time.py cfg2 > gpurun_out/ncu_run.log 2>&1
tail -5 gpurun_out/ncu_plain.log gpurun_out/ncu_run.log
