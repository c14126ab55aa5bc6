#!/bin/bash
# developer sweep: build kernel-configuration variants into gpurun-visible variants/ (git-ignored *.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() { # name NWARPS CAP RCAP STAGES MINB U
  nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -DPDLP_NWARPS=$2 -DPDLP_CAP=$3 -DPDLP_RCAP=$4 -DPDLP_STAGES=$5 -DPDLP_MINB=$6 -DPDLP_U=$7 $8 $9 \
    -Xptxas -v -o variants/lib_$1.so torchpdlp_b200/csrc/pdlp_b200.cu 2> variants/$1.ptxas.log
  echo "$1: $(grep -A1 'DualOpIdLb0ELb0' variants/$1.ptxas.log | grep -o 'Used [0-9]* registers') spills: $(grep -A2 'DualOpIdLb0ELb0' variants/$1.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
}
for cfg in "$@"; do build $cfg & done
wait
