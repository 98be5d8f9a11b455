#!/bin/bash
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_gemm.py tests/test_gpu_fp32_mode.py -m gpu -q -x 2>&1 | tail -40 > $out/r2d_pytest_new.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_base.py tests/test_gpu_parity_large.py tests/test_collator_cfg1.py -m gpu -q -x 2>&1 | tail -15 > $out/r2d_pytest_model.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > $out/r2d_bench_cfg3.json 2> $out/r2d_bench_cfg3.err
tail -c 300 $out/r2d_bench_cfg3.err
timeout 600 python tools/dropout_cost.py cfg4 > $out/r2d_dropout_cost.log 2>&1
LAKO_DECODE_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2d_launches_decode.csv \
    python tools/prof_decode.py 8 20 250 3 > /dev/null 2>&1
LAKO_DECODE_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'xattn_decode_kernel' \
    --launch-skip 30 -c 3 -f -o /tmp/r2d_decode python tools/prof_decode.py > $out/r2d_ncu_decode.log 2>&1
ncu -i /tmp/r2d_decode.ncu-rep --page raw --csv > $out/r2d_ncu_full_xattn_decode.csv 2>/dev/null
tail -6 $out/r2d_pytest_new.log; tail -4 $out/r2d_pytest_model.log; cat $out/r2d_dropout_cost.log | tail -3
