#!/bin/bash
# round-end confirmation on one B200: GPU parity suite, the default bench line, ncu captures, per-format bench lines
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
python bench.py 2> gpurun_out/final_bench_rfc5424.err | tail -1 > gpurun_out/final_bench_rfc5424.json; cut -c1-400 gpurun_out/final_bench_rfc5424.json
ncu --set full --clock-control none --import-source on -k regex:parse_kernel -s 3 -c 1 -o gpurun_out/prof_r1i python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r1i.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_r1i.log 2>&1
python bench.py --format ltsv --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_ltsv.json
python bench.py --format ltsv --ltsv-typed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_ltsv_typed.json
python bench.py --format gelf --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_gelf.json
ncu --set full --clock-control none --import-source on -k regex:parse_kernel -s 3 -c 1 -o gpurun_out/prof_r1i_ltsv python bench.py --format ltsv --lines 500000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r1i_ltsv.log 2>&1
python bench.py --format mixed --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_mixed.json
python bench.py --split --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_split.json
ls -la gpurun_out | tail -15
