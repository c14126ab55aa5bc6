#!/bin/bash
# developer run: GPU parity tests (default library), then timing of every variant in variants/ on a config
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CFG=${1:-cfg2}
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_auto.log 2>&1; echo "tests auto rc=$?" >> gpurun_out/r2_tests_auto.log
  PDLP_PANEL_MIN_NNZ=0 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_panel_small.log 2>&1; echo "tests panel rc=$?" >> gpurun_out/r2_tests_panel_small.log
  tail -n 4 gpurun_out/r2_tests_auto.log gpurun_out/r2_tests_panel_small.log
fi
export PDLP_TIMING_ONLY=1
for f in variants/lib_*.so; do
  n=$(basename $f .so)
  PDLP_B200_LIB=$PWD/$f timeout 200 python scripts/dev_time.py $CFG > gpurun_out/r2_sweep_$n.log 2>&1
  echo "== $n: $(grep -E 'graph=True|kernels:' gpurun_out/r2_sweep_$n.log | tr '\n' ' ')"
done
