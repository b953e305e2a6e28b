#!/bin/bash
# round 2: GELF member-parallel with lock-step line pass and member validation
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 900 python -m pytest tests/test_gpu_gelf.py tests/test_gpu_encode.py tests/test_gpu_pipeline.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; tail -5 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --format gelf --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2n_bench_gelf.json
timeout 600 python bench.py --format mixed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2n_bench_mixed.json
for f in gelf mixed; do python -c "import json; d=json.load(open('gpurun_out/r2n_bench_$f.json')); print('$f', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_gelf_kernel -s 3 -c 1 -o gpurun_out/prof_r2n_gelf python bench.py --format gelf --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2n_gelf.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "gelf or mixed" > gpurun_out/r2n_pytest_full.log 2>&1; tail -3 gpurun_out/r2n_pytest_full.log
ls gpurun_out | tail -6
