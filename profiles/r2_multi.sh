#!/bin/bash
# round 2, N GPUs of one box (gpurun --gpus N): the driver's launch line for bench.py, the mixed C5 stream, and the
# two-real-devices fan-out test.  usage: bash profiles/r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_gpu_fullsize.py::test_multi_device_fanout tests/test_gpu_pipeline.py::test_multi_context_fanout -x -q -m gpu > gpurun_out/r2m_pytest_n$N.log 2>&1; tail -3 gpurun_out/r2m_pytest_n$N.log
for k in 1 2 4 8; do
  if [ $k -le $N ]; then
    if [ $k -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2m_err_$k.log | tail -1 > gpurun_out/r2m_bench_rfc5424_n$k.json
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $k --steps 20 --warmup 3 2>gpurun_out/r2m_err_$k.log | tail -1 > gpurun_out/r2m_bench_rfc5424_n$k.json
    fi
    python -c "import json; d=json.load(open('gpurun_out/r2m_bench_rfc5424_n$k.json')); print('N=$k', 'value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'e2e_record', d.get('e2e_record',{}).get('value'))" || tail -5 gpurun_out/r2m_err_$k.log
  fi
done
if [ $N -ge 2 ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --format mixed --steps 10 --warmup 3 2>gpurun_out/r2m_err_mixed.log | tail -1 > gpurun_out/r2m_bench_mixed_n$N.json
  python -c "import json; d=json.load(open('gpurun_out/r2m_bench_mixed_n$N.json')); print('mixed N=$N', d['value'], d['per_gpu_lines_per_s'], d['e2e']['value'])" || tail -5 gpurun_out/r2m_err_mixed.log
fi
ls gpurun_out | tail -12
