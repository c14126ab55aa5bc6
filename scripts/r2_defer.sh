#!/bin/bash
# developer run, 1 GPU: K1's grid reduction deferred into K2's finalize -- full suite, then config-2 timings with and without
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_defer_tests.log 2>&1
echo "== suite rc=$?: $(tail -n 1 gpurun_out/r2_defer_tests.log)"
grep -E "^E |Error" gpurun_out/r2_defer_tests.log | head -20
for v in "PDLP_NO_DEFER=1" "PDLP_X=1" "PDLP_NO_DEFER=1" "PDLP_X=1"; do
  echo "== cfg2 $v"
  env $v PDLP_TIMING_ONLY=1 timeout 300 python scripts/dev_time.py cfg2 2>&1 | grep -E "graph=True|kernels:|Error|error"
done
