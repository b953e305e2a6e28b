"""CPU model of the shipped RFC3164 walker: r3164_parse_lockstep run with 32 host threads per warp (tests/emu, libfg_emu_warp.so),
counting the lock-step iterations (warp votes) per source line.  A model of where the iterations go, NOT a measurement.
    python profiles/lockstep_votes_r3164.py > profiles/r2s_lockstep_votes_rfc3164.txt"""
import sys, time
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
for d in ('', 'oracle', 'tests', 'tests/emu'):
    sys.path.insert(0, str(REPO / d))
import numpy as np
import flowgger_b200 as fb, emu as E
E.build_warp(True)
n=4096
data,offs=fb.generate(3, 3164, n, bad_frac=0.005)
g,go,info=E.r3164_decode_dump(fb,data,offs,2026,warp=True)
sv=E.r3164_site_votes()
src=open(REPO / 'flowgger_b200/csrc/fg_rfc3164.cuh').read().splitlines()
tot=sum(sv.values())
print(f"# r3164_parse_lockstep on {n} generated lines (seed 3164, the bench mix) = {n//32} warps: {tot} rendezvous in total = {tot/(n/32):.0f} per warp = {tot/n:.1f} per line; mismatches {info['vote_mismatches']}")
print("# a rendezvous = one lock-step iteration of the loop at that source line (the __syncwarp phase boundaries are not listed)")
for l,c in sorted(sv.items(), key=lambda kv:-kv[1]):
    print(f"{c:8d} {100*c/tot:5.1f}%  {c/(n/32):7.1f}/warp  fg_rfc3164.cuh:{l}  {src[l-1].strip()[:100] if l<len(src) else ''}")
