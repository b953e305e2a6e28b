#!/bin/bash
# where does the e2e time go?  (1) normal  (2) results not copied back (H2D + kernels only)
for v in "" "FG_DEBUG_SKIP_D2H=1"; do
  env $v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 > /tmp/b.json
  python - "$v" <<'PY'
import json, sys
d = json.loads(open('/tmp/b.json').read())
print(sys.argv[1] or "normal", d["e2e"]["value"], d["e2e"]["gb_per_s"], d["e2e"]["kernel_ms_per_step"])
PY
done
