#!/bin/bash
# developer run, 1 GPU: full GPU suite, then the N=1 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_suite_final.log 2>&1
echo "== suite rc=$?: $(tail -n 1 gpurun_out/r2_suite_final.log)"
timeout 900 python bench.py --steps ${STEPS:-200} --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "== bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","us_per_iteration","iteration_roofline_frac")})
print("e2e", d["e2e"]["value"], d["time_to_tolerance"]); print("roofline", d["roofline"]["primal"], d["roofline"]["dual"]); print("cfg4", d.get("cfg4"))
PY
tail -n 3 gpurun_out/r2_bench_n1.err
