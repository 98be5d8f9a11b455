#!/bin/bash
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -30 > $out/r2e_pytest_new.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_base.py tests/test_gpu_parity_large.py tests/test_collator_cfg1.py tests/test_gpu_fp32_mode.py -m gpu -q 2>&1 | tail -15 > $out/r2e_pytest_model.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > $out/r2e_bench_cfg3.json 2> $out/r2e_bench_cfg3.err
tail -c 300 $out/r2e_bench_cfg3.err
LAKO_DECODE_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2e_launches_decode.csv \
    python tools/prof_decode.py 8 20 250 3 > /dev/null 2>&1
tail -6 $out/r2e_pytest_new.log; tail -6 $out/r2e_pytest_model.log
