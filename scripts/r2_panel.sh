#!/bin/bash
# developer run: GPU parity tests + timing of the panel layout with several sorting windows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_auto.log 2>&1; echo "tests auto rc=$?" >> gpurun_out/r2_tests_auto.log
PDLP_PANEL_MIN_NNZ=0 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_panel_small.log 2>&1; echo "tests panel rc=$?" >> gpurun_out/r2_tests_panel_small.log
tail -n 4 gpurun_out/r2_tests_auto.log gpurun_out/r2_tests_panel_small.log
export PDLP_TIMING_ONLY=1 PDLP_TRACE_SETUP=1
for sg in 128 512 2048 8192; do
  PDLP_PANEL_SIGMA=$sg timeout 200 python scripts/dev_time.py cfg2 > gpurun_out/r2_panel_sigma$sg.log 2>&1
  echo "== sigma>=$sg: $(grep -E 'K\^T panels:' gpurun_out/r2_panel_sigma$sg.log | cut -c1-160)"
  echo "   $(grep -E 'graph=True|kernels:' gpurun_out/r2_panel_sigma$sg.log | tr '\n' ' ')"
done
PDLP_LAYOUT=sell timeout 200 python scripts/dev_time.py cfg2 2>&1 | grep -E 'graph=True|kernels:' | tr '\n' ' '
