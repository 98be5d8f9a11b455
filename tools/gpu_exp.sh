#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step']); print(d.get('generate') or d.get('decode_step') or {k:v for k,v in d.items() if 'decode' in k or 'generate' in k})"
