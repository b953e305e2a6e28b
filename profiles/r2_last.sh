#!/bin/bash
# round 2: last check of the shipped build — every GPU test file, then the driver's smoke()
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
: > gpurun_out/r2y_pytest.log
for t in tests/test_gpu_*.py; do
  timeout 300 python -m pytest $t -x -q -m gpu >> gpurun_out/r2y_pytest.log 2>&1; rc=$?
  echo "$t rc=$rc $(tail -1 gpurun_out/r2y_pytest.log)"
  if [ $rc -ne 0 ] && [ $rc -ne 5 ]; then echo "parity failed or hung in $t: stopping"; exit 1; fi
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee gpurun_out/r2y_smoke.log
