#!/bin/bash
# staged (TMA tile in shared memory) vs unstaged (lines read through L1) for the long-line formats
for fmt in ltsv gelf; do for v in "" "FG_FORCE_STAGE=1"; do
  env $v python bench.py --format $fmt --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 > /tmp/b.json
  python - $fmt "$v" <<'PY'
import json, sys
d = json.loads(open('/tmp/b.json').read())
print(sys.argv[1], sys.argv[2] or "default(unstaged)", d["ms_per_step"], d["roofline"]["frac"], d["value"])
PY
done; done
