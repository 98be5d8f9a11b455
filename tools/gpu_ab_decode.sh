#!/bin/bash
# same-box A/B of the decode step (expects _lib/liblako_b200_old.so next to the new one)
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
for f in old "" old "" ; do
lib=lako_b200/_lib/liblako_b200${f:+_$f}.so
echo "== $lib"
LAKO_B200_LIB=$PWD/$lib timeout 600 python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value'],2), d['generate']['decode_step'])"
done
