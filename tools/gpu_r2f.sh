#!/bin/bash
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "enc_attn" 2>&1 | tail -25 > $out/r2f_pytest_attn.log
tail -4 $out/r2f_pytest_attn.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropout.py tests/test_gpu_parity_base.py tests/test_gpu_parity_large.py -m gpu -q 2>&1 | tail -15 > $out/r2f_pytest_model.log
tail -4 $out/r2f_pytest_model.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $out/r2f_bench_cfg4.json 2> $out/r2f_bench_cfg4.err
tail -c 300 $out/r2f_bench_cfg4.err
LAKO_ATTN_FWD=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > $out/r2f_bench_cfg4_attn1.json 2> $out/r2f_bench_cfg4_attn1.err
LAKO_DECODE_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'xattn_decode_kernel' \
    --launch-skip 30 -c 2 -f -o /tmp/r2f_decode python tools/prof_decode.py > $out/r2f_ncu_decode.log 2>&1
ncu -i /tmp/r2f_decode.ncu-rep --page raw --csv > $out/r2f_ncu_full_xattn_decode.csv 2>/dev/null
