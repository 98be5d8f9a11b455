#!/bin/bash
for b in 1 32; do
timeout 900 python bench.py --workload cfg3 --batch $b --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_cfg3_b$b.json 2> gpurun_out/r2w_bench_cfg3_b$b.err
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench_cfg3_b$b.json')); print($b, d['value'], d['ms_per_step'], d.get('generate',{}).get('decode_step'), d['e2e']['value'])" || tail -3 gpurun_out/r2w_bench_cfg3_b$b.err
done
