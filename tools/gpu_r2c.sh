#!/bin/bash
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_fp32_mode.py tests/test_gpu_elementwise.py tests/test_collator_cfg1.py tests/test_cpu_host.py -m gpu -q -rA -s 2>&1 | tail -150 > $out/r2c_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fp32_mode.py --deselect tests/test_gpu_decode.py 2>&1 | tail -30 > $out/r2c_pytest_all.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > $out/r2c_bench_cfg3.json 2> $out/r2c_bench_cfg3.err
tail -c 300 $out/r2c_bench_cfg3.err
timeout 900 python bench.py --steps 10 --warmup 3 > $out/r2c_bench_cfg4.json 2> $out/r2c_bench_cfg4.err
tail -c 300 $out/r2c_bench_cfg4.err
LAKO_DECODE_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2c_launches_decode.csv \
    python tools/prof_decode.py 8 20 250 3 > /dev/null 2>&1
grep -E "passed|failed|PASSED|FAILED|XFAIL|Error|error|fp32 mode|tf32x3" $out/r2c_pytest_new.log | tail -60
tail -5 $out/r2c_pytest_all.log
