#!/bin/bash
# A/B builds of the RFC5424 kernel shape (CTA width / residency); run on the CPU box, the .so files travel with gpurun.
#   usage: bash profiles/variants.sh          -> flowgger_b200/lib_v_<lines>_<minb>/libflowgger_cuda.so
set -e
cd "$(dirname "$0")/.."
for v in "64 14 104" "64 15 102"; do
  set -- $v
  d=flowgger_b200/lib_v_$1_$2${4:+_${4#-DFG_R5_}}
  mkdir -p $d
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --fmad=false \
    -DFG_R5_LINES=$1 -DFG_R5_MINB=$2 -DFG_TILE_SLACK_PCT=$3 $4 -shared -o $d/libflowgger_cuda.so flowgger_b200/csrc/*.cu -I include -Xptxas -v 2>&1 | grep -A2 parse5424_kernel | grep -E "registers|spill" | sed "s/^/[$1 x $2] /"
done
