#!/bin/bash
# developer run: full ncu capture of the two iteration kernels on config 4 (final kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PDLP_TIMING_ONLY=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 420 -c 2 \
    -o gpurun_out/r2_cfg4 -f python scripts/dev_time.py cfg4 > gpurun_out/ncu4.log 2>&1
echo "ncu rc=$?"; tail -n 4 gpurun_out/ncu4.log
