#!/bin/bash
# round 2: LTSV part-parallel, rows placed by popcount rank, parse_ts classes on separate warps: parity, bench, ncu
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 900 python -m pytest tests/test_gpu_ltsv.py tests/test_gpu_pipeline.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; tail -5 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --format ltsv --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2j_bench_ltsv.json
timeout 600 python bench.py --format ltsv --ltsv-typed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2j_bench_ltsv_typed.json
for f in ltsv ltsv_typed; do python -c "import json; d=json.load(open('gpurun_out/r2j_bench_$f.json')); print('$f', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_ltsv_kernel -s 3 -c 1 -o gpurun_out/prof_r2j_ltsv python bench.py --format ltsv --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2j_ltsv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_ltsv_kernel -s 3 -c 1 -o gpurun_out/prof_r2j_ltsv_typed python bench.py --format ltsv --ltsv-typed --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2j_ltsv_typed.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "ltsv or mixed" > gpurun_out/r2j_pytest_full.log 2>&1; tail -3 gpurun_out/r2j_pytest_full.log
ls gpurun_out | tail -6
