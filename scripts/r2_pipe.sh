#!/bin/bash
# developer run, 1 GPU: pipelined-solve tests, full GPU suite, e2e probe, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solve.py -m gpu -x -q -k "pipelined or time_limit or kkt_limit" > gpurun_out/r2_pipe_tests.log 2>&1
echo "== pipelined tests rc=$?: $(tail -n 1 gpurun_out/r2_pipe_tests.log)"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_suite_pipe.log 2>&1
echo "== suite rc=$?: $(tail -n 1 gpurun_out/r2_suite_pipe.log)"
PDLP_TRACE_SETUP=1 timeout 600 python scripts/dev_e2e_ingest.py > gpurun_out/r2_e2e_pipe.log 2>&1
echo "== e2e probe rc=$?"; grep "PDLP_HOST_INGEST" gpurun_out/r2_e2e_pipe.log
PDLP_NO_PIPELINE=1 PDLP_NO_PREBUILD=1 timeout 600 python scripts/dev_e2e_ingest.py > gpurun_out/r2_e2e_nopipe.log 2>&1
echo "== e2e probe (no pipeline, no prebuild) rc=$?"; grep "PDLP_HOST_INGEST" gpurun_out/r2_e2e_nopipe.log
timeout 900 python bench.py --steps ${STEPS:-200} --warmup 5 > gpurun_out/r2_bench_n1c.json 2> gpurun_out/r2_bench_n1c.err
echo "== bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1c.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","us_per_iteration","iteration_roofline_frac")})
print("e2e", d["e2e"]["value"], d["time_to_tolerance"]); print("roofline", d["roofline"]["primal"], d["roofline"]["dual"])
PY
tail -n 3 gpurun_out/r2_bench_n1c.err
