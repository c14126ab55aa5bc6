#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PDLP_TIMING_ONLY=1 PDLP_TRACE_SETUP=1
for MB in 0 40 32 22 16 11; do
  PDLP_WINDOW_MB=$MB timeout 300 python scripts/dev_time.py cfg4 > gpurun_out/r2_cfg4_win$MB.log 2>&1
  echo "== window $MB MB: $(grep -E 'graph=True|kernels:' gpurun_out/r2_cfg4_win$MB.log | tr '\n' ' ') $(grep -oE 'K in [0-9]+ column windows' gpurun_out/r2_cfg4_win$MB.log)"
done
