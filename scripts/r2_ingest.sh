#!/bin/bash
# developer run, 1 GPU: ingest tests, full GPU suite, e2e probe, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "ingest or power" > gpurun_out/r2_ingest_tests.log 2>&1
echo "== ingest tests rc=$?: $(tail -n 1 gpurun_out/r2_ingest_tests.log)"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_suite_ingest.log 2>&1
echo "== suite rc=$?: $(tail -n 1 gpurun_out/r2_suite_ingest.log)"
PDLP_TRACE_SETUP=1 timeout 600 python scripts/dev_e2e_ingest.py > gpurun_out/r2_e2e_ingest.log 2>&1
echo "== e2e probe rc=$?"; grep "PDLP_HOST_INGEST" gpurun_out/r2_e2e_ingest.log
timeout 900 python bench.py --steps ${STEPS:-200} --warmup 5 > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err
echo "== bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1b.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","us_per_iteration","iteration_roofline_frac")})
print("e2e", d["e2e"]["value"], d["time_to_tolerance"]); print("roofline", d["roofline"]["primal"], d["roofline"]["dual"])
PY
tail -n 3 gpurun_out/r2_bench_n1b.err
