#!/bin/bash
# developer run: launch list (durations) + full capture of the panel / combine kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PDLP_TIMING_ONLY=1
timeout 300 python scripts/dev_time.py cfg2 > gpurun_out/ncu_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 12 --csv --log-file gpurun_out/r2_launches_panel.csv python scripts/dev_time.py cfg2 > gpurun_out/ncu_run0.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'panel_kernel|combine_kernel' -s 60 -c 2 \
    -o gpurun_out/r2_prof_panel2 -f python scripts/dev_time.py cfg2 > gpurun_out/ncu_run.log 2>&1
grep -E "panel_kernel|combine_kernel|stream_kernel" gpurun_out/r2_launches_panel.csv | awk -F'","' '{print $5, $NF}' | head -12
