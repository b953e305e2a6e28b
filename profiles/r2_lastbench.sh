#!/bin/bash
# round 2: the bench lines of the shipped build (512 Ki-line chunks, roofline.traffic from profiles/traffic.json)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 3 --encode --split 2>/dev/null | tail -1 > gpurun_out/r2y_bench_rfc5424.json; cut -c1-300 gpurun_out/r2y_bench_rfc5424.json
timeout 200 python bench.py --format ltsv --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2y_bench_ltsv.json
timeout 200 python bench.py --format gelf --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2y_bench_gelf.json
for f in rfc5424 ltsv gelf; do python -c "import json; d=json.load(open('gpurun_out/r2y_bench_$f.json')); r=d['roofline']; print('$f', d['value'], d['kernel_ms'], r['frac'], r['traffic'], d['e2e']['value'], d.get('gpu_launches'))"; done
