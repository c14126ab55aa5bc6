#!/bin/bash
# developer sweep: kernel-configuration variants (variants/lib_*.so, scripts/build_variants.sh) on config 2
cd "$(dirname "$0")/.."
export PDLP_TIMING_ONLY=1
echo "== default"; python scripts/dev_time.py cfg2 2>&1 | grep -E "graph=True|kernels:"
for v in variants/lib_*.so; do
  echo "== $v"; PDLP_B200_LIB=$PWD/$v python scripts/dev_time.py cfg2 2>&1 | grep -E "graph=True|kernels:|rror"
done
