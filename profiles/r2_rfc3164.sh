#!/bin/bash
# RFC3164 (N3) on the device: the new GPU tests, then one bench line.  Run with what was left of the round's GPU budget.
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_z_rfc3164.py -x -q > gpurun_out/r3_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3_pytest.log
tail -25 gpurun_out/r3_pytest.log
timeout 150 python bench.py --format rfc3164 --lines 4000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
echo "bench exit $?"
cut -c1-1500 gpurun_out/r3_bench.json
tail -5 gpurun_out/r3_bench.err
