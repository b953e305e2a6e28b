#!/bin/bash
# round 2: compute-sanitizer over smoke() (every shipped kernel on 4096 lines per format): memcheck, then racecheck of the
# shared-memory phases
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_sanitize_memcheck.log; tail -4 gpurun_out/r2_sanitize_memcheck.log
timeout 260 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 10 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/r2_sanitize_racecheck.log; tail -6 gpurun_out/r2_sanitize_racecheck.log
