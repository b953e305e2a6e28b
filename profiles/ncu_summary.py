#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into text: headline metrics + hottest SASS by executed count.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct", "gpu__dram_throughput.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warps_issue_stalled", "launch__grid_size",
        "launch__shared_mem_per_block", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_lsu"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) for w in want) and not h.endswith(("max_rate",)):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
ia, isrc, iavg = h.index("Instructions Executed"), h.index("Source"), h.index("Avg. Threads Executed")
isamp = h.index("# Samples")
tot = sum(int(r[ia]) for r in data); tsamp = sum(int(r[isamp]) for r in data)
print(f"\ntotal warp instructions {tot}; samples {tsamp}")
idx = sorted(range(len(data)), key=lambda k: -int(data[k][ia]))[:top]
print("idx  executed  avg_threads samples  sass")
for k in sorted(idx):
    r = data[k]
    print(f"{k:5d} {int(r[ia]):10d} {r[iavg]:>5s} {int(r[isamp]):6d}  {r[isrc].strip()[:90]}")
