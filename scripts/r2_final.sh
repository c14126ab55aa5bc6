#!/bin/bash
# developer run, 1 GPU: full GPU suite, smoke, bench line (driver's arguments and the default), ncu launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_suite_final.log 2>&1
echo "== suite rc=$?: $(tail -n 1 gpurun_out/r2_suite_final.log)"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1
echo "== smoke rc=$?: $(tail -n 1 gpurun_out/r2_smoke.log)"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_s20.json 2> gpurun_out/r2_bench_n1_s20.err
echo "== bench(20) rc=$?"
timeout 600 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
echo "== bench(200) rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r2_bench_n1_s20.json", "gpurun_out/r2_bench_n1_final.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print({k:d[k] for k in ("value","ms_per_step","us_per_iteration","iteration_roofline_frac")})
        print("e2e", d["e2e"]["value"], d["time_to_tolerance"]); print("roofline", d["roofline"]["frac"], d["roofline"]["primal"], d["roofline"]["dual"], d["roofline"].get("loop_ms_per_iteration"))
        print("cfg4", d.get("cfg4"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -n 3 gpurun_out/r2_bench_n1_final.err
