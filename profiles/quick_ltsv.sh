#!/bin/bash
# LTSV parity tests + kernel-resident time, plain and typed (prints ms_per_step, roofline frac, value, e2e)
python -m pytest tests/test_gpu_ltsv.py tests/test_gpu_split.py -x -q 2>&1 | tail -2
for extra in "" "--ltsv-typed"; do
  python bench.py --format ltsv $extra --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 > /tmp/b.json
  python - <<'PY'
import json
d = json.loads(open('/tmp/b.json').read())
print(d["ms_per_step"], d["roofline"]["frac"], d["value"], d["e2e"]["value"])
PY
done
