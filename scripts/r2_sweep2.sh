#!/bin/bash
# developer run: timing of every variant in variants/ for several layouts (no tests)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PDLP_TIMING_ONLY=1 PDLP_TRACE_SETUP=1
for f in variants/lib_*.so; do
  n=$(basename $f .so)
  for L in ${LAYOUTS:-panel sell}; do
    for SG in ${SIGMAS:-128}; do
      PDLP_PANEL_SIGMA=$SG PDLP_LAYOUT=$L PDLP_B200_LIB=$PWD/$f timeout 200 python scripts/dev_time.py ${CFG:-cfg2} > gpurun_out/r2_sw_${n}_${L}_$SG.log 2>&1
      echo "== $n $L sigma>=$SG: $(grep -E 'graph=True|kernels:' gpurun_out/r2_sw_${n}_${L}_$SG.log | tr '\n' ' ') $(grep -oE 'sigma [0-9]+, [0-9]+ pieces, [0-9.]+ MB packed, [0-9]+ ring stages' gpurun_out/r2_sw_${n}_${L}_$SG.log)"
      [ "$L" = "sell" ] && break
    done
  done
done
