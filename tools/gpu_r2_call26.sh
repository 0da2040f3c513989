#!/bin/bash
# round 2, call 26: compute-sanitizer (memcheck, racecheck) over every batch path of the final tree
mkdir -p gpurun_out
timeout 700 compute-sanitizer --tool memcheck python tools/sanitize_paths.py 60001 > gpurun_out/s26_memcheck.log 2>&1; tail -16 gpurun_out/s26_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_paths.py 20001 > gpurun_out/s26_racecheck.log 2>&1; tail -16 gpurun_out/s26_racecheck.log
