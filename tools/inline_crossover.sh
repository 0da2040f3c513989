for n in 1e8 2e8 4e8 8e8; do echo "== n=$n"; python tools/c3_inline.py $n 5e7 2>&1 | grep -E "mutated': False" | grep -E "'inline': 0|'blocks_per_sm': (3|4|5)," ; done
