#!/bin/bash
# round 2, final: the whole GPU parity suite, the bench line of every config, ncu captures + launch lists of the shipped kernels
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so  # the exact build the profiles below belong to (tools/ncu_by_line.py)
: > gpurun_out/r2z_pytest.log
for t in tests/test_gpu_*.py; do   # file by file, so that a hang costs one timeout and stops the run
  timeout 420 python -m pytest $t -x -q -m gpu >> gpurun_out/r2z_pytest.log 2>&1; rc=$?
  echo "$t rc=$rc $(tail -1 gpurun_out/r2z_pytest.log)"
  if [ $rc -ne 0 ] && [ $rc -ne 5 ]; then echo "parity failed or hung in $t: stopping"; exit 1; fi
done
timeout 300 python bench.py --steps 20 --warmup 3 --encode --split 2> gpurun_out/r2z_bench.err | tail -1 > gpurun_out/r2z_bench_rfc5424.json; cut -c1-400 gpurun_out/r2z_bench_rfc5424.json; tail -3 gpurun_out/r2z_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2z_bench_reference.json; cut -c1-300 gpurun_out/r2z_bench_reference.json
timeout 300 python bench.py --format ltsv 2>/dev/null | tail -1 > gpurun_out/r2z_bench_ltsv.json
timeout 300 python bench.py --format ltsv --ltsv-typed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2z_bench_ltsv_typed.json
timeout 300 python bench.py --format gelf 2>/dev/null | tail -1 > gpurun_out/r2z_bench_gelf.json
timeout 300 python bench.py --format mixed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2z_bench_mixed.json
for f in ltsv ltsv_typed gelf mixed; do python -c "import json; d=json.load(open('gpurun_out/r2z_bench_$f.json')); print('$f', d['value'], d['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_r2z_$3 ${@:4} > gpurun_out/ncu_r2z_$3.log 2>&1; }
cap parse5424_kernel 3 parse5424 python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1
cap post5424_kernel 3 post5424 python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1
cap parse_ltsv_kernel 3 ltsv python bench.py --format ltsv --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1
cap parse_gelf_kernel 3 gelf python bench.py --format gelf --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1
cap gelf_write_kernel 0 gelfw python profiles/enc_probe.py 1000000
cap gelf_size_kernel 0 gelfs python profiles/enc_probe.py 1000000
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2z.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --encode > gpurun_out/launches_r2z.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2z_ltsv.csv python bench.py --format ltsv --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2z_gelf.csv python bench.py --format gelf --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
ls -la gpurun_out | tail -20
