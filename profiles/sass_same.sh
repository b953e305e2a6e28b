#!/bin/bash
# Are the kernels of the last profiled commit and of the working tree the same SASS?  (Run on the CPU box.)
#   usage: bash profiles/sass_same.sh <commit>
# Used after the RFC3164 work (which touched fg_kernels.cuh / fg_kernels.cu / fg_abi.cu): parse5424 / parse_ltsv / parse_gelf /
# gelf_encode of 31f1fe2 and of the tree are instruction-for-instruction identical, so the ncu captures and
# profiles/traffic.json of the r2z run still describe the shipped RFC5424 / LTSV / GELF / encoder kernels.
set -e
cd "$(dirname "$0")/.."
old=$(mktemp -d)
git archive "$1" flowgger_b200/csrc include | tar -x -C "$old"
flags="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -cubin"
for k in fg_parse5424 fg_parse_ltsv fg_parse_gelf fg_gelf_encode fg_split $( [ -f "$old/flowgger_b200/csrc/fg_parse3164.cu" ] && echo fg_parse3164 ); do
  (cd "$old" && /usr/local/cuda/bin/nvcc $flags -o $k.old.cubin flowgger_b200/csrc/$k.cu -I include 2>/dev/null)
  /usr/local/cuda/bin/nvcc $flags -o "$old/$k.new.cubin" flowgger_b200/csrc/$k.cu -I include 2>/dev/null
  for f in old new; do
    /usr/local/cuda/bin/cuobjdump -sass "$old/$k.$f.cubin" | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > "$old/$k.$f.sass"
  done
  if cmp -s "$old/$k.old.sass" "$old/$k.new.sass"; then echo "$k: identical ($(wc -l < "$old/$k.new.sass") instructions)"; else echo "$k: DIFFERENT"; fi
done
