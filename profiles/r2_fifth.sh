#!/bin/bash
# round 2: LTSV on the bitmap pipeline (parse_ltsv_kernel) + lock-step / length-sorted GELF encoder: parity, bench, ncu
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 900 python -m pytest tests/test_gpu_ltsv.py tests/test_gpu_encode.py tests/test_gpu_pipeline.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
timeout 300 python profiles/enc_probe.py 1000000 2>&1 | tail -3
timeout 600 python bench.py --format ltsv --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2e_bench_ltsv.json
timeout 600 python bench.py --format ltsv --ltsv-typed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2e_bench_ltsv_typed.json
timeout 600 python bench.py --format mixed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2e_bench_mixed.json
for f in ltsv ltsv_typed mixed; do python -c "import json; d=json.load(open('gpurun_out/r2e_bench_$f.json')); print('$f', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --encode 2>/dev/null | tail -1 > gpurun_out/r2e_bench_rfc5424.json; python -c "import json; d=json.load(open('gpurun_out/r2e_bench_rfc5424.json')); print('step_ms', d['kernel_ms'], 'e2e', d['e2e']['value'], 'encode', d.get('encode_e2e'))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_ltsv_kernel -s 3 -c 1 -o gpurun_out/prof_r2e_ltsv python bench.py --format ltsv --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2e_ltsv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gelf_write_kernel -c 1 -o gpurun_out/prof_r2e_gelfw python profiles/enc_probe.py 1000000 > gpurun_out/ncu_r2e_gelfw.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r2e_pytest_full.log 2>&1; tail -3 gpurun_out/r2e_pytest_full.log
ls gpurun_out | tail -8
