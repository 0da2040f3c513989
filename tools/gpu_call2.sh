#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1v}
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python tools/layouts.py 1e8 5e7 plain,packed3,packed4 15 > $OUT/${TAG}_layouts_c2.log 2>&1; echo "layouts c2 rc=$?"
timeout 600 python tools/layouts.py 2e8 5e7 plain,packed3,packed4 15 > $OUT/${TAG}_layouts_200m.log 2>&1; echo "layouts 200m rc=$?"
timeout 600 python tools/layouts.py 4e8 5e7 packed3,packed4 15 > $OUT/${TAG}_layouts_400m.log 2>&1; echo "layouts 400m rc=$?"
timeout 900 python tools/layouts.py 3.1e9 2.5e8 packed3,packed4 15 > $OUT/${TAG}_layouts_c3.log 2>&1; echo "layouts c3 rc=$?"
timeout 600 python tools/e2e_chunks.py 1e8 5e7 > $OUT/${TAG}_e2e_chunks.log 2>&1; echo "e2e chunks rc=$?"; cat $OUT/${TAG}_e2e_chunks.log
grep -h "Gq_per_s" $OUT/${TAG}_layouts_*.log | python -c "
import sys,ast
for l in sys.stdin:
    r=ast.literal_eval(l)
    print(r['genome_bp'],r['layout'],'refill',r['refill'],'mut',int(r['mutated_half']),'bps',r['blocks_per_sm'],r['Gq_per_s'],r['same_results'])
"
