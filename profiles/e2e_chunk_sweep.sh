#!/bin/bash
# e2e (fg_decode_batch, pinned host buffers) vs pipeline chunk size
for c in 65536 262144 1048576 4194304; do
  FG_CHUNK_LINES=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 > /tmp/b.json
  python - "$c" <<'PY'
import json, sys
d = json.loads(open('/tmp/b.json').read())
print(sys.argv[1], d["e2e"]["value"], d["e2e"]["gb_per_s"], d["e2e"]["kernel_ms_per_step"])
PY
done
