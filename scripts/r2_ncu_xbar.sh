#!/bin/bash
# developer run: full ncu capture of the two iteration kernels on config 4 with K2 on the dense xbar vector
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PDLP_TIMING_ONLY=1 PDLP_XBAR=${PDLP_XBAR:-1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 420 -c 2 \
    -o gpurun_out/r2_prof_cfg4_xbar -f python scripts/dev_time.py cfg4 > gpurun_out/ncu_run_xbar.log 2>&1
echo "ncu rc=$?"; tail -n 5 gpurun_out/ncu_run_xbar.log
