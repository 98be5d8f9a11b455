#!/bin/bash
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "enc_attn" 2>&1 | tail -25 > $out/r2h_pytest_attn.log
tail -4 $out/r2h_pytest_attn.log
LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_trace.so timeout 300 python tools/attn_fwd_trace.py > $out/r2h_attn_trace.log 2>&1
cat $out/r2h_attn_trace.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropout.py tests/test_gpu_parity_base.py -m gpu -q 2>&1 | tail -8 > $out/r2h_pytest_model.log
tail -4 $out/r2h_pytest_model.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > $out/r2h_bench_cfg4.json 2> $out/r2h_bench_cfg4.err
tail -c 300 $out/r2h_bench_cfg4.err
python -c "
import json; d=json.load(open('$out/r2h_bench_cfg4.json')); print(round(d['value'],2), round(d['ms_per_step'],2), d['kernels']['enc_attn_fwd'], d['clocks'])"
