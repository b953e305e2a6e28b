#!/bin/bash
# round 2: GELF member-parallel (word-level line pass, dense number pass) + LTSV typed values grouped by type
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 900 python -m pytest tests/test_gpu_gelf.py tests/test_gpu_ltsv.py tests/test_gpu_pipeline.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r2k_pytest.log 2>&1; tail -5 gpurun_out/r2k_pytest.log
timeout 600 python bench.py --format gelf --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2k_bench_gelf.json
timeout 600 python bench.py --format ltsv --ltsv-typed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2k_bench_ltsv_typed.json
timeout 600 python bench.py --format mixed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2k_bench_mixed.json
for f in gelf ltsv_typed mixed; do python -c "import json; d=json.load(open('gpurun_out/r2k_bench_$f.json')); print('$f', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_gelf_kernel -s 3 -c 1 -o gpurun_out/prof_r2k_gelf python bench.py --format gelf --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2k_gelf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_ltsv_kernel -s 3 -c 1 -o gpurun_out/prof_r2k_ltsv_typed python bench.py --format ltsv --ltsv-typed --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2k_ltsv_typed.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "gelf or mixed" > gpurun_out/r2k_pytest_full.log 2>&1; tail -3 gpurun_out/r2k_pytest_full.log
ls gpurun_out | tail -6
