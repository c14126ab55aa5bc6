#!/bin/bash
# developer run on N GPUs: hardware parity tests of the row-partitioned engine, then the N-GPU bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
if [ "$2" != "notests" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_multi_tests_$N.log 2>&1
  echo "== multi tests rc=$?: $(tail -n 1 gpurun_out/r2_multi_tests_$N.log)"
fi
PDLP_BENCH_TRACE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps ${STEPS:-50} --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "== bench rc=$?"
tail -c 3000 gpurun_out/r2_bench_n$N.json
grep -E "cfg5|Error|error" gpurun_out/r2_bench_n$N.err | tail -n 8
