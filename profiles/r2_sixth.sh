#!/bin/bash
# round 2: LTSV walker with parked values (no per-key work inside the part loop), CTA-width A/B, ncu
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 900 python -m pytest tests/test_gpu_ltsv.py -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; tail -5 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --format ltsv --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2f_bench_ltsv.json
timeout 600 python bench.py --format ltsv --ltsv-typed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2f_bench_ltsv_typed.json
for f in ltsv ltsv_typed; do python -c "import json; d=json.load(open('gpurun_out/r2f_bench_$f.json')); print('$f', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
for d in flowgger_b200/lib_v_lt*; do
  echo "== $d"; FG_VARIANT_DIR=$d timeout 300 python bench.py --format ltsv --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('step_ms', d['kernel_ms'], 'frac', d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/r2f_variants.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_ltsv_kernel -s 3 -c 1 -o gpurun_out/prof_r2f_ltsv python bench.py --format ltsv --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2f_ltsv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse_ltsv_kernel -s 3 -c 1 -o gpurun_out/prof_r2f_ltsv_typed python bench.py --format ltsv --ltsv-typed --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2f_ltsv_typed.log 2>&1
ls gpurun_out | tail -6
