#!/bin/bash
# RFC3164 lock-step walker: GPU tests + the three-data-set timing (the last seconds of the round's GPU budget)
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_z_rfc3164.py -x -q > gpurun_out/r3c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3c_pytest.log; tail -4 gpurun_out/r3c_pytest.log
timeout 30 python profiles/quick_r3164.py 2000000 lockstep 2>&1 | tee gpurun_out/r3c_timing.txt
