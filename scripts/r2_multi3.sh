#!/bin/bash
# developer run, N GPUs: two-rank parity tests (dense xbar, windows, exchanges), then the bench line with the config-5 record
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "${KEXPR:-xbar or windows or exchanges}" > gpurun_out/r2_multi_tests_$N.log 2>&1
echo "== multi tests rc=$?: $(tail -n 1 gpurun_out/r2_multi_tests_$N.log)"
PDLP_TRACE_SETUP=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps ${STEPS:-100} --warmup 3 > gpurun_out/r2_bench_n${N}_xbar.json 2> gpurun_out/r2_bench_n${N}_xbar.err
echo "== bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_xbar.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","us_per_iteration","iteration_roofline_frac")}, d["parity"]["ok"])
print(json.dumps(d.get("cfg5")))
PY
grep -E "column windows|xbar|ranges" gpurun_out/r2_bench_n${N}_xbar.err | sort | uniq -c | head -20
tail -n 5 gpurun_out/r2_bench_n${N}_xbar.err
