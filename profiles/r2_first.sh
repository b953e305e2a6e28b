#!/bin/bash
# round 2, first GPU contact of the bitmap + bit-walk RFC5424 kernel: parity, bench line, launch list, one full ncu capture
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so  # the exact build the profiles below belong to (tools/ncu_by_line.py)
timeout 900 python -m pytest tests/test_gpu_rfc5424.py tests/test_gpu_pipeline.py tests/test_gpu_split.py tests/test_gpu_encode.py -x -q -m gpu --deselect tests/test_gpu_rfc5424.py::test_full_size_batch_parity > gpurun_out/r2a_pytest.log 2>&1; tail -15 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/r2a_bench.err | tail -1 > gpurun_out/r2a_bench_rfc5424.json; cut -c1-1500 gpurun_out/r2a_bench_rfc5424.json; tail -3 gpurun_out/r2a_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse5424_kernel -s 3 -c 1 -o gpurun_out/prof_r2a python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post5424_kernel -s 3 -c 1 -o gpurun_out/prof_r2a_post python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2a_post.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_r2a.log 2>&1
ls -la gpurun_out | tail -8
