#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counts with nvdisasm line info: warp instructions, thread instructions and
stall samples per SOURCE LINE of the kernel (build with -lineinfo).

usage: python tools/ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [lines_processed]
If lines_processed is given, counts are also shown per processed log line."""
import csv, io, re, subprocess, sys, tempfile, os, collections

rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
per = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ia, ie, it, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
counts = [(int(r[ie]), int(r[it]), int(r[isamp])) for r in rows[2:] if len(r) > isamp]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname not in dis:
        continue
    # every .text section whose name holds the substring is a candidate (template instantiations): the one whose
    # instruction count matches the report wins
    cur, infn, out, cands = ("?", 0), False, [], []
    for ln in dis.splitlines():
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            if out:
                cands.append(out)
            out = []
            infn = kname in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            out.append(cur)
    if out:
        cands.append(out)
    for out in cands:
        if lines is None or abs(len(out) - len(counts)) < abs(len(lines) - len(counts)):
            lines = out
assert lines is not None, "kernel not found"
if len(lines) != len(counts):
    print(f"warning: {len(lines)} disassembled instructions vs {len(counts)} in the report", file=sys.stderr)
agg = collections.OrderedDict()
for (f, l), (e, t, s) in zip(lines, counts):
    a = agg.setdefault((f, l), [0, 0, 0])
    a[0] += e; a[1] += t; a[2] += s
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"total warp instr {tot}" + (f" = {tot / per:.1f} per line" if per else "") + f"; samples {tots}")
src_cache = {}
def src(f, l):
    if f not in src_cache:
        for root in ("flowgger_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                src_cache[f] = open(p).read().splitlines(); break
        else:
            src_cache[f] = []
    s = src_cache[f]
    return s[l - 1].strip()[:100] if 0 < l <= len(s) else ""
print(f"{'file:line':28s} {'warp-instr':>11s} {'%':>5s} {'/line':>6s} {'thr/instr':>9s} {'samples%':>8s}  source")
for (f, l), (e, t, s) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if e * 200 < tot and s * 200 < tots:
        continue
    print(f"{f + ':' + str(l):28s} {e:11d} {100 * e / tot:5.1f} {e / per if per else 0:6.2f} {t / max(e, 1):9.1f} {100 * s / max(tots, 1):8.1f}  {src(f, l)}")

# totals per source file and per function-sized block of 25 lines: where the long tail of small lines sits
by_file = collections.OrderedDict()
for (f, l), (e, t, s_) in agg.items():
    a = by_file.setdefault((f, l // 25 * 25), [0, 0])
    a[0] += e; a[1] += t
print("\nby 25-line block (>= 1 %):")
for (f, l), (e, t) in sorted(by_file.items()):
    if e * 100 >= tot:
        print(f"{f + ':' + str(l) + '+':28s} {e:11d} {100 * e / tot:5.1f} {e / per if per else 0:6.2f} {t / max(e, 1):9.1f}")
