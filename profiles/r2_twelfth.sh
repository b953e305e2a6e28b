#!/bin/bash
# round 2: GELF member-parallel (alignment fix) + encoder 4-bytes-per-iteration
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 900 python -m pytest tests/test_gpu_gelf.py tests/test_gpu_encode.py tests/test_gpu_pipeline.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r2l_pytest.log 2>&1; tail -5 gpurun_out/r2l_pytest.log
timeout 600 python bench.py --format gelf --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2l_bench_gelf.json
timeout 600 python bench.py --format mixed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2l_bench_mixed.json
for f in gelf mixed; do python -c "import json; d=json.load(open('gpurun_out/r2l_bench_$f.json')); print('$f', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
timeout 300 python profiles/enc_probe.py 1000000 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --encode 2>/dev/null | tail -1 > gpurun_out/r2l_bench_rfc5424.json; python -c "import json; d=json.load(open('gpurun_out/r2l_bench_rfc5424.json')); print('step_ms', d['kernel_ms'], 'e2e', d['e2e']['value'], 'encode', d.get('encode_e2e'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_gelf_kernel -s 3 -c 1 -o gpurun_out/prof_r2l_gelf python bench.py --format gelf --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2l_gelf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gelf_write_kernel -c 1 -o gpurun_out/prof_r2l_gelfw python profiles/enc_probe.py 1000000 > gpurun_out/ncu_r2l_gelfw.log 2>&1
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "gelf or mixed" > gpurun_out/r2l_pytest_full.log 2>&1; tail -3 gpurun_out/r2l_pytest_full.log
ls gpurun_out | tail -6
