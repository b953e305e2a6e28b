#!/bin/bash
# round 2: GELF lock-step member-parallel kernel + encoder uniform push: parity first (abort on a failure or a hang), then bench / ncu
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 240 python -m pytest tests/test_gpu_gelf.py -x -q -m gpu > gpurun_out/r2o_pytest.log 2>&1; rc=$?; tail -5 gpurun_out/r2o_pytest.log
if [ $rc -ne 0 ]; then echo "GELF parity failed or hung (rc=$rc): stopping"; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_encode.py tests/test_gpu_pipeline.py tests/test_gpu_split.py -x -q -m gpu > gpurun_out/r2o_pytest2.log 2>&1; rc=$?; tail -3 gpurun_out/r2o_pytest2.log
if [ $rc -ne 0 ]; then echo "encode / pipeline parity failed or hung (rc=$rc): stopping"; exit 1; fi
timeout 300 python bench.py --format gelf --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2o_bench_gelf.json
python -c "import json; d=json.load(open('gpurun_out/r2o_bench_gelf.json')); print('gelf', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"
timeout 300 python profiles/enc_probe.py 1000000 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:parse_gelf_kernel -s 3 -c 1 -o gpurun_out/prof_r2o_gelf python bench.py --format gelf --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2o_gelf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gelf_write_kernel -c 1 -o gpurun_out/prof_r2o_gelfw python profiles/enc_probe.py 1000000 > gpurun_out/ncu_r2o_gelfw.log 2>&1
ls gpurun_out | tail -5
