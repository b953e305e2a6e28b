#!/bin/bash
# round 2: 512 Ki-line chunks (e2e), LTSV / GELF with 128-thread CTAs (A/B against the shipped 256)
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 300 python -m pytest tests/test_gpu_ltsv.py tests/test_gpu_gelf.py tests/test_gpu_pipeline.py tests/test_gpu_encode.py -x -q -m gpu > gpurun_out/r2p_pytest.log 2>&1; rc=$?; tail -2 gpurun_out/r2p_pytest.log
if [ $rc -ne 0 ]; then echo "parity failed or hung (rc=$rc): stopping"; exit 1; fi
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 --encode 2>/dev/null | tail -1 > gpurun_out/r2p_bench_rfc5424.json; python -c "import json; d=json.load(open('gpurun_out/r2p_bench_rfc5424.json')); print('rfc5424 step_ms', d['kernel_ms'], 'e2e', d['e2e']['value'], 'e2e_record', d['e2e_record']['value'], 'encode', d['encode_e2e']['value'])"
for f in ltsv gelf; do timeout 200 python bench.py --format $f --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('$f 256 threads: step_ms', d['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"; done
export FG_VARIANT_DIR=flowgger_b200/lib_v_t128
timeout 200 python -m pytest tests/test_gpu_ltsv.py tests/test_gpu_gelf.py -x -q -m gpu > gpurun_out/r2p_pytest_t128.log 2>&1; rc=$?; tail -2 gpurun_out/r2p_pytest_t128.log
if [ $rc -ne 0 ]; then echo "128-thread variant: parity failed or hung (rc=$rc): stopping"; exit 1; fi
for f in ltsv gelf; do timeout 200 python bench.py --format $f --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('$f 128 threads: step_ms', d['kernel_ms'], 'frac', d['roofline']['frac'])"; done
