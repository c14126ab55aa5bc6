#!/bin/bash
# developer run, 1 GPU: host-ingest tests, e2e probe with K^T planned on a helper thread (and without), cfg4 ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ingest or host" > gpurun_out/r2_ingest_tests.log 2>&1
echo "== ingest tests rc=$?: $(tail -n 1 gpurun_out/r2_ingest_tests.log)"
PDLP_TRACE_SETUP=1 timeout 300 python scripts/dev_e2e_ingest.py > gpurun_out/r2_e2e_thread.log 2>&1
echo "== e2e rc=$?"; grep "PDLP_HOST_INGEST=1" gpurun_out/r2_e2e_thread.log; grep -E "plan K|wait:|engine_create total" gpurun_out/r2_e2e_thread.log | tail -6
PDLP_NO_PLAN_THREAD=1 timeout 300 python scripts/dev_e2e_ingest.py 2>&1 | grep "PDLP_HOST_INGEST=1"
bash scripts/r2_ncu_cfg4.sh
