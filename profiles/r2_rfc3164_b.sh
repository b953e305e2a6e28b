#!/bin/bash
# RFC3164: one ncu capture of parse3164_kernel + timings of the base build and three CTA shapes on three data sets
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:parse3164_kernel -s 2 -c 1 -o gpurun_out/prof_r3a python profiles/quick_r3164.py 1000000 ncu mix > gpurun_out/ncu_r3a.log 2>&1
tail -2 gpurun_out/ncu_r3a.log
timeout 60 python profiles/quick_r3164.py 2000000 base 2>&1 | tee gpurun_out/r3a_variants.txt
for v in 128_8 32_32 64_8; do
  FG_VARIANT_DIR=flowgger_b200/lib_r3_$v timeout 60 python profiles/quick_r3164.py 2000000 $v 2>&1 | tee -a gpurun_out/r3a_variants.txt
done
