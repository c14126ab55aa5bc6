#!/bin/bash
cd "$(dirname "$0")/.."
for f in variants/lib_*.so; do
  echo "=== $f"
  PDLP_B200_LIB=$PWD/$f timeout 120 python scripts/dev_time.py ${1:-cfg2} 2>&1 | grep -E "graph=|solve:|Error|error|spmv|power"
done
