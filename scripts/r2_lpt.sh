#!/bin/bash
# developer run, 1 GPU: planner deals slices out longest-first (LPT) for wide sorting windows -- tests, config 4 timings, balance probe on config 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py -m gpu -x -q -k "ragged or cfg4 or xbar or reproducible or spmv" > gpurun_out/r2_lpt_tests.log 2>&1
echo "== tests rc=$?: $(tail -n 1 gpurun_out/r2_lpt_tests.log)"
for v in "PDLP_XBAR=1" "PDLP_XBAR=1 PDLP_NO_LPT=1" "PDLP_XBAR=0 PDLP_WINDOW_MB=0" "PDLP_XBAR=0"; do
  echo "== cfg4 $v"
  env $v PDLP_TIMING_ONLY=1 PDLP_TRACE_SETUP=1 timeout 300 python scripts/dev_time.py cfg4 2>&1 | grep -E "graph=True|kernels:|column windows|ranges|Error|error" 
done
for c in cfg2 cfg2even; do
  echo "== $c"
  PDLP_TIMING_ONLY=1 timeout 300 python scripts/dev_time.py $c 2>&1 | grep -E "nnz=|graph=True|kernels:|Error|error"
done
