#!/bin/bash
# occupancy experiment: FG_MINB selects a kernel instantiation compiled for another CTAs/SM bound
# usage: minb_sweep.sh <format> <minb...>
fmt=$1; shift
for mb in "$@"; do
  FG_MINB=$mb python bench.py --format $fmt --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 > /tmp/b.json
  python - $fmt $mb <<'PY'
import json, sys
d = json.loads(open('/tmp/b.json').read())
print(sys.argv[1], "minb", sys.argv[2], d["ms_per_step"], d["roofline"]["frac"], d["value"])
PY
done
