"""SIMT cost model of the RFC5424 structured-data walk on the C2 synthetic lines (CPU only, no GPU needed).

The parse kernel walks SD in lock step: one OUTER iteration handles, per lane, one sd_id or one name="value" pair, and the
byte scans inside are warp-uniform loops, so a warp pays max-over-lanes of (a) the number of outer iterations and (b) the
16-byte blocks each scan touches.  This script measures, on generated lines, how much of that maximum is imbalance and what
other lane<->line mappings would pay.  It backs the round-2 plan in DESIGN.md section 9.

    python tools/simt_model.py [n_lines]
"""
import re
import sys

import numpy as np

sys.path.insert(0, ".")
from flowgger_b200 import native  # noqa: E402

PAIR = re.compile(rb'([!-<>-~]+)="((?:[^"\\]|\\.)*)"')


def blocks16(start: int, end: int) -> int:
    """16-byte aligned blocks a scan from `start` up to the stop byte at `end` touches"""
    return (end >> 4) - (start >> 4) + 1


def line_steps(line: bytes, off: int):
    """per outer iteration: (id blocks, name 4-byte words, value blocks) — offsets taken relative to the buffer (alignment)"""
    parts = line.split(b" ", 6)
    if len(parts) < 7 or not parts[6].startswith(b"["):
        return []
    sd = parts[6]
    base = off + len(line) - len(sd)
    steps = []
    i = 0
    while i < len(sd) and sd[i:i + 1] == b"[":
        j = sd.find(b" ", i)
        k = sd.find(b"]", i)
        if j < 0 or (0 <= k < j):
            j = k
        steps.append((blocks16(base + i + 1, base + j), 0, 0))
        pos = j
        while True:
            m = PAIR.match(sd, pos + 1) if sd[pos:pos + 1] == b" " else None
            if not m:
                break
            ns, ne = m.span(1)
            vs, ve = m.span(2)
            steps.append((0, (base + ne >> 2) - (base + ns >> 2) + 1, blocks16(base + vs, base + ve)))
            pos = m.end()
        close = sd.find(b"]", pos)
        if close < 0:
            break
        steps.append((0, 0, 0))  # the ']' classification step
        i = close + 1
    return steps


def warp_cost(groups):
    """groups: list (one per lane) of step lists; cost = sum over outer iterations of max-over-lanes per scan"""
    n = max((len(g) for g in groups), default=0)
    outer = n
    a = c = d = 0
    for t in range(n):
        col = [g[t] for g in groups if t < len(g)]
        a += max(x[0] for x in col)
        c += max(x[1] for x in col)
        d += max(x[2] for x in col)
    return outer, a, c, d


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 1024
    data, offsets = native.generate(0, 1234, n, bad_frac=0.0)[:2]
    buf = data.tobytes()
    lines = [(buf[offsets[i]:offsets[i + 1]], int(offsets[i])) for i in range(n)]
    steps = [line_steps(l, o) for l, o in lines]
    per_line = np.array([len(s) for s in steps])
    print(f"lines {n}, with SD {np.mean(per_line > 0):.2f}, outer steps per line mean {per_line.mean():.2f} max {per_line.max()}")
    # instruction weights per loop iteration, from the r1i SASS: outer fixed ~70, id block ~40, name word ~22, value block ~60
    W = dict(outer=70, a=40, c=22, d=60)

    def report(name, warps, lines_per_warp):
        tot = np.zeros(4)
        for g in warps:
            tot += np.array(warp_cost(g))
        nl = len(warps) * lines_per_warp
        instr = (tot[0] * W["outer"] + tot[1] * W["a"] + tot[2] * W["c"] + tot[3] * W["d"]) / nl
        print(f"{name:58s} outer/line {tot[0] / nl:5.2f}  id {tot[1] / nl:5.2f}  name {tot[2] / nl:5.2f}  value {tot[3] / nl:5.2f}  SD warp-instr/line {instr:6.1f}")

    # (1) today: 32 consecutive lines per warp, one line per lane
    report("1 line per lane (today)", [steps[i:i + 32] for i in range(0, n, 32)], 32)
    # (2) K consecutive lines per lane, walked back to back (a lane concatenates the steps of its K lines)
    for K in (2, 4, 8):
        warps = []
        for w in range(0, n - 32 * K + 1, 32 * K):
            warps.append([sum((steps[w + l * K + k] for k in range(K)), []) for l in range(32)])
        report(f"{K} lines per lane, back to back", warps, 32 * K)
    # (3) a CTA of 128 lines deals its SD work out by greedy longest-first bin packing over 128 lanes' worth of 4 warps
    for K in (4,):
        warps = []
        for w in range(0, n - 32 * K + 1, 32 * K):
            chunk = sorted(steps[w:w + 32 * K], key=len, reverse=True)
            lanes = [[] for _ in range(32)]
            for s in chunk:
                min(lanes, key=len).extend(s)
            warps.append(lanes)
        report(f"{K} lines per lane, longest-first dealt to the shortest lane", warps, 32 * K)
    # (4) lines sorted by their number of outer steps inside a CTA, then one line per lane (warps of similar lines);
    #     "SD lines compacted" is the coarse version of this that was tried in round 1 (slower: hand-off > gain)
    for cta in (64, 128, 256):
        warps = []
        for w in range(0, n - cta + 1, cta):
            chunk = sorted(steps[w:w + cta], key=len)
            warps += [chunk[k:k + 32] for k in range(0, cta, 32)]
        report(f"sorted by #steps inside a {cta}-line CTA, 1 line per lane", warps, 32)
    warps = []
    for w in range(0, n - 127, 128):
        chunk = sorted(steps[w:w + 128], key=lambda s: len(s) > 0)
        warps += [chunk[k:k + 32] for k in range(0, 128, 32)]
    report("SD-bearing lines compacted inside a 128-line CTA (tried)", warps, 32)
    # lower bound: perfect balance (mean over lanes instead of max)
    tot = np.zeros(4)
    for s in steps:
        for x in s:
            tot += np.array([1, x[0], x[1], x[2]])
    instr = (tot[0] * W["outer"] + tot[1] * W["a"] + tot[2] * W["c"] + tot[3] * W["d"]) / n / 32
    print(f"{'perfect balance (sum over lines / 32 lanes)':58s} outer/line {tot[0] / n / 32:5.2f}  SD warp-instr/line {instr:6.1f}")


if __name__ == "__main__":
    main()
