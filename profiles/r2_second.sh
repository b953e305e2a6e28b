#!/bin/bash
# round 2: parity suite, bench lines, kernel-shape A/B (profiles/variants.sh builds), ncu captures of both RFC5424 kernels
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so  # the exact build the profiles below belong to (tools/ncu_by_line.py)
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_rfc5424.py::test_full_size_batch_parity --deselect tests/test_gpu_fullsize.py::test_ltsv_10m_lines --deselect tests/test_gpu_fullsize.py::test_gelf_10m_lines --deselect tests/test_gpu_fullsize.py::test_timestamp_bits_vs_python_mini_oracle > gpurun_out/r2b_pytest.log 2>&1; tail -12 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --encode 2> gpurun_out/r2b_bench.err | tail -1 > gpurun_out/r2b_bench_rfc5424.json; cut -c1-2200 gpurun_out/r2b_bench_rfc5424.json; tail -3 gpurun_out/r2b_bench.err
for d in flowgger_b200/lib_v_*; do
  echo "== $d"; FG_VARIANT_DIR=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('kernel_ms', d['kernel_ms'], 'frac', d['roofline']['frac'])"
done 2>&1 | tee gpurun_out/r2b_variants.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:parse5424_kernel -s 3 -c 1 -o gpurun_out/prof_r2b python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post5424_kernel -s 3 -c 1 -o gpurun_out/prof_r2b_post python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2b_post.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/launches_r2b.log 2>&1
ls -la gpurun_out | tail -12
