#!/bin/bash
# developer run, 1 GPU: dense-xbar K2 (DualBarOp) -- parity tests, then timings on configs 4 and 2 against the pair kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solve.py -m gpu -x -q -k "xbar or windows or cfg4 or sharded" > gpurun_out/r2_xbar_tests.log 2>&1
echo "== xbar tests rc=$?: $(tail -n 1 gpurun_out/r2_xbar_tests.log)"
grep -E "Error|error|assert" gpurun_out/r2_xbar_tests.log | head -20
for v in "PDLP_XBAR=0" "PDLP_XBAR=1" "PDLP_XBAR=1 PDLP_WINDOW_MB=20" "PDLP_XBAR=0 PDLP_WINDOW_MB=0"; do
  echo "== cfg4 $v"
  env $v PDLP_TIMING_ONLY=1 PDLP_TRACE_SETUP=1 timeout 300 python scripts/dev_time.py cfg4 2>&1 | grep -E "graph=True|kernels:|column windows|xbar|Error|error" 
done
for v in "PDLP_XBAR=0" "PDLP_XBAR=1"; do
  echo "== cfg2 $v"
  env $v PDLP_TIMING_ONLY=1 timeout 300 python scripts/dev_time.py cfg2 2>&1 | grep -E "graph=True|kernels:|Error|error"
done
