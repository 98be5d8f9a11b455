#!/bin/bash
# same-box A/B of two library builds: tools/ab.sh <workload> <steps> <tagA> <tagB> ...  ("cur" = the in-tree build)
w=$1; steps=$2; shift 2
for rep in 1 2; do
  for tag in "$@"; do
    if [ "$tag" = cur ]; then unset LAKO_B200_LIB; else export LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_$tag.so; fi
    timeout 300 python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); c=d['clocks']
print('$tag', '$w', round(d['value'],1), round(d['ms_per_step'],3), c['sm_mhz'], c.get('power_w'), c['reasons'])"
  done
done
