#!/bin/bash
# GPU session: parity tests, layout A/B incl. lane refill at c2 / c3, bench c2 + c3.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1v}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/${TAG}_box.txt 2>&1; nproc >> $OUT/${TAG}_box.txt; free -g >> $OUT/${TAG}_box.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python tools/layouts.py 1e8 5e7 plain,packed3,packed4 15 > $OUT/${TAG}_layouts_c2.log 2>&1; echo "layouts c2 rc=$?"
timeout 900 python tools/layouts.py 3.1e9 2.5e8 inline,packed3,packed4 15 > $OUT/${TAG}_layouts_c3.log 2>&1; echo "layouts c3 rc=$?"
grep -h "Gq_per_s" $OUT/${TAG}_layouts_*.log | python -c "
import sys,ast
for l in sys.stdin:
    r=ast.literal_eval(l)
    print(r['genome_bp'],r['layout'],'refill',r['refill'],'mut',int(r['mutated_half']),'bps',r['blocks_per_sm'],r['Gq_per_s'],r['same_results'])
"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 900 python bench.py --workload c3 --steps 5 > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err; echo "bench c3 rc=$?"; cat $OUT/${TAG}_bench_c3.json
