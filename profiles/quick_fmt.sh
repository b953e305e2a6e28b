#!/bin/bash
# usage: quick_fmt.sh <format> <test files...> : parity tests + kernel-resident time (ms_per_step, roofline frac, value, e2e)
fmt=$1; shift
python -m pytest "$@" -x -q 2>&1 | tail -2
python bench.py --format $fmt --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>&1 | tail -1 > /tmp/b.json
python - <<'PY'
import json
d = json.loads(open('/tmp/b.json').read())
print(d["ms_per_step"], d["roofline"]["frac"], d["value"], d["e2e"]["value"])
PY
