#!/bin/bash
# A/B builds of the RFC3164 kernel shape; run on the CPU box, the .so files travel with gpurun.
set -e
cd "$(dirname "$0")/.."
for v in "128 8" "32 32" "64 8"; do
  set -- $v
  d=flowgger_b200/lib_r3_$1_$2
  mkdir -p $d
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --fmad=false \
    -DFG_R3_LINES=$1 -DFG_R3_MINB=$2 -shared -o $d/libflowgger_cuda.so flowgger_b200/csrc/*.cu -I include -Xptxas -v 2>&1 | grep -A2 parse3164_kernel | grep -E "registers|spill" | sed "s/^/[$1 x $2] /"
done
