"""RFC3164 kernel timing on three data sets (generated mix / one regular line repeated / irregular spacing only): separates the
cost of the common path from the re-join path and from divergence.  usage: python profiles/quick_r3164.py [lines] [tag]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import flowgger_b200 as fb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
tag = sys.argv[2] if len(sys.argv) > 2 else "base"
sets = {}
sets["mix"] = fb.generate(fb.FMT_RFC3164, 3164, n, bad_frac=0.005)
line = b"<134>Aug  6 11:15:24 web-01.example.org nginx[2231]: GET /api/v1/items 200 13 ms from 10.2.3.4 session opened for user id 4711 ok"
reg = np.frombuffer(line * n, dtype=np.uint8).copy()
sets["regular"] = (reg, (np.arange(n + 1, dtype=np.int64) * len(line)).astype(np.int32))
irr = line.replace(b"items 200", b"items  200")
ir = np.frombuffer(irr * n, dtype=np.uint8).copy()
sets["irregular"] = (ir, (np.arange(n + 1, dtype=np.int64) * len(irr)).astype(np.int32))
only = sys.argv[3] if len(sys.argv) > 3 else None
for name, (data, offs) in sets.items():
    if only and name != only:
        continue
    dec = fb.BatchDecoder(fb.FMT_RFC3164, max_batch_bytes=int(offs[-1]) + (1 << 20), max_batch_lines=n, rfc3164_year=2026)
    dec.upload(data, offs)
    for _ in range(3):
        dec.parse_resident()
    ms = dec.parse_resident_many(10) / 10
    res = dec.download()
    print(f"{tag} {name}: {ms:.3f} ms / {n} lines = {n / ms / 1e6:.2f} G lines/s, {int(offs[-1]) / ms / 1e6:.1f} GB/s, "
          f"errors {int((res.status != 0).sum())}, arena {len(res.arena)}", flush=True)
    dec.close()
