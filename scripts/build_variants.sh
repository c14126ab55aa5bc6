#!/bin/bash
# developer sweep: build kernel-configuration variants into variants/ (git-ignored, shipped by gpurun)
# usage: build_variants.sh "name W KB U CTAS OVERSUB [extra -D flags]" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() {
  name=$1; W=$2; KB=$3; U=$4; CTAS=$5; OV=$6; shift 6
  nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -DPDLP_W=$W -DPDLP_KB=$KB -DPDLP_U=$U -DPDLP_CTAS=$CTAS -DPDLP_MINB=$CTAS -DPDLP_OVERSUB=$OV "$@" \
    -Xptxas -v -o variants/lib_$name.so torchpdlp_b200/csrc/pdlp_b200.cu 2> variants/$name.ptxas.log
  echo "$name: $(grep -A1 'DualOpIdLb0ELb0' variants/$name.ptxas.log | grep -o 'Used [0-9]* registers') $(grep -A2 'DualOpIdLb0ELb0' variants/$name.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
}
for cfg in "$@"; do build $cfg & done
wait
