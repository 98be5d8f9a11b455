#!/bin/bash
# same-box comparison of decode-kernel variants (every _lib/liblako_b200*.so) on the cfg3 decode step
for r in 1 2; do for lib in lako_b200/_lib/liblako_b200*.so; do
echo "== $lib"
LAKO_B200_LIB=$PWD/$lib timeout 600 python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value'],2), d['generate']['decode_step']['ms_per_token'])"
done; done
