#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_solve.py -m gpu -x -q -k "sharded" > gpurun_out/r2_shard1.log 2>&1
echo "== single-rank sharded tests rc=$?: $(tail -n 1 gpurun_out/r2_shard1.log)"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_multi_tests_$N.log 2>&1
echo "== multi tests rc=$?: $(tail -n 1 gpurun_out/r2_multi_tests_$N.log)"
PDLP_BENCH_CFG5=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/r2_bench_n${N}_fused.json 2> gpurun_out/r2_bench_n${N}_fused.err
echo "== bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_fused.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","us_per_iteration","iteration_roofline_frac")}, d["parity"]["ok"], d["e2e"]["what"])
PY
tail -n 5 gpurun_out/r2_bench_n${N}_fused.err
