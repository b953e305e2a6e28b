#!/bin/bash
# round 2: encoder launch list, lock-step unescape A/B, e2e chunk sweep
mkdir -p gpurun_out
cp flowgger_b200/lib/libflowgger_cuda.so gpurun_out/lib_used.so
timeout 600 python -m pytest tests/test_gpu_rfc5424.py tests/test_gpu_encode.py tests/test_gpu_pipeline.py -x -q -m gpu --deselect tests/test_gpu_rfc5424.py::test_full_size_batch_parity > gpurun_out/r2d_pytest.log 2>&1; tail -3 gpurun_out/r2d_pytest.log
timeout 300 python profiles/enc_probe.py 1000000 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2d_enc.csv python profiles/enc_probe.py 1000000 > gpurun_out/launches_r2d_enc.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_r2d_enc.csv')))
hdr=None; agg=collections.defaultdict(list)
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: agg[d['Kernel Name'][:60]].append(float(d['Metric Value'].replace(',','')))
        except: pass
for k,v in agg.items(): print(k, len(v), round(sum(v)/len(v)/1000,1),'us avg', round(sum(v)/1000,1), 'us total')
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -1 > gpurun_out/r2d_bench_rfc5424.json; python -c "import json; d=json.load(open('gpurun_out/r2d_bench_rfc5424.json')); print('step_ms', d['kernel_ms'], 'dominant_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
for c in 131072 524288 1048576; do FG_CHUNK_LINES=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('chunk $c e2e', d['e2e']['value'])"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:post5424_kernel -s 3 -c 1 -o gpurun_out/prof_r2d_post python bench.py --lines 1000000 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_r2d_post.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gelf_size_kernel -c 1 -o gpurun_out/prof_r2d_gelfs python profiles/enc_probe.py 1000000 > gpurun_out/ncu_r2d_gelfs.log 2>&1
ls gpurun_out | tail -5
