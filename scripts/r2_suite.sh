#!/bin/bash
# developer run: the GPU suite under every layout of the iteration kernels, then config-4 timing per layout
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for L in sell run panel; do
  PDLP_LAYOUT=$L PDLP_PANEL_MIN_NNZ=0 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_suite_$L.log 2>&1
  echo "== suite layout=$L rc=$?: $(tail -n 1 gpurun_out/r2_suite_$L.log)"
done
export PDLP_TIMING_ONLY=1
for L in sell run; do
  PDLP_LAYOUT=$L timeout 400 python scripts/dev_time.py cfg4 > gpurun_out/r2_cfg4_$L.log 2>&1
  echo "== cfg4 $L: $(grep -E 'graph=True|kernels:' gpurun_out/r2_cfg4_$L.log | tr '\n' ' ')"
done
