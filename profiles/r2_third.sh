#!/bin/bash
# round 2: the whole GPU parity suite (full-size tests included), the bench line of every config, ncu captures of the shipped kernels
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so  # the exact build the profiles below belong to (tools/ncu_by_line.py)
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; tail -6 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --encode --split 2> gpurun_out/r2c_bench.err | tail -1 > gpurun_out/r2c_bench_rfc5424.json; cut -c1-600 gpurun_out/r2c_bench_rfc5424.json; tail -3 gpurun_out/r2c_bench.err
for d in flowgger_b200/lib_v_*; do
  echo "== $d"; FG_VARIANT_DIR=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('step_ms', d['kernel_ms'], 'dominant_ms', d['roofline'].get('kernel_ms'), 'frac', d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/r2c_variants.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2c_bench_reference.json; cut -c1-300 gpurun_out/r2c_bench_reference.json
timeout 600 python bench.py --format ltsv --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c_bench_ltsv.json
timeout 600 python bench.py --format ltsv --ltsv-typed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c_bench_ltsv_typed.json
timeout 600 python bench.py --format gelf --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c_bench_gelf.json
timeout 600 python bench.py --format mixed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2c_bench_mixed.json
for f in ltsv gelf mixed; do python -c "import json; d=json.load(open('gpurun_out/r2c_bench_$f.json')); print('$f', d['value'], d['roofline']['frac'], d['e2e']['value'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse5424_kernel -s 3 -c 1 -o gpurun_out/prof_r2c python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post5424_kernel -s 3 -c 1 -o gpurun_out/prof_r2c_post python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2c_post.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gelf_write_kernel -c 1 -o gpurun_out/prof_r2c_gelfw python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 --encode > gpurun_out/ncu_r2c_gelfw.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --encode > gpurun_out/launches_r2c.log 2>&1
ls -la gpurun_out | tail -14
