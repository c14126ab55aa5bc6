#!/bin/bash
# round-2 developer run: GPU parity tests on the new layouts + timing of the three layouts on config 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_auto.log 2>&1; echo "tests auto rc=$?" >> gpurun_out/r2_tests_auto.log
PDLP_PANEL_MIN_NNZ=0 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_panel_small.log 2>&1; echo "tests panel rc=$?" >> gpurun_out/r2_tests_panel_small.log
for L in sell run panel; do
  PDLP_LAYOUT=$L PDLP_TRACE_SETUP=1 timeout 300 python scripts/dev_time.py cfg2 > gpurun_out/r2_time_$L.log 2>&1
done
tail -3 gpurun_out/r2_tests_auto.log gpurun_out/r2_tests_panel_small.log
grep -E "graph=|kernels:|solve:" gpurun_out/r2_time_*.log
