#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -X faulthandler bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/dbg_c1.json 2> gpurun_out/dbg_c1.log; echo "rc=$?"; tail -30 gpurun_out/dbg_c1.log
