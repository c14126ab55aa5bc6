#!/bin/bash
# developer sweep: build kernel-configuration variants into variants/ (git-ignored, shipped by gpurun)
# usage: build_variants.sh "name -DFLAG=.. -DFLAG=.." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() {
  name=$1; shift
  nvcc -shared -Xcompiler -fPIC -Xcompiler -pthread -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 "$@" \
    -Xptxas -v -o variants/lib_$name.so torchpdlp_b200/csrc/pdlp_b200.cu 2> variants/$name.ptxas.log
  echo "$name: built"
}
for cfg in "$@"; do build $cfg & done
wait
