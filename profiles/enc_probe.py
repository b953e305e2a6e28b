"""Probe of the fused decode + GELF encode path (fg_decode_encode_gelf): a few calls on a small batch, for an ncu launch list."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import flowgger_b200 as fb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
data, offs = fb.generate(fb.FMT_RFC5424, 5424, n, mean_len=169.2, bad_frac=0.005, nthreads=16)
dec = fb.BatchDecoder(fb.FMT_RFC5424, max_batch_bytes=int(offs[-1]) + (1 << 20), max_batch_lines=n)
hb = dec.host_alloc(int(offs[-1]))
ho = dec.host_alloc(offs.nbytes, dtype=np.int32)
hb[:] = data
ho[:] = offs
for k in range(3):
    buf, o, st, kms = dec.decode_encode_gelf(hb, ho, copy=False)
    print("call", k, "kernel_ms", kms, "json bytes", int(o[-1]))
dec.close()
