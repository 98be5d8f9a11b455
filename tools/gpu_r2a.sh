#!/bin/bash
# round-2 first measurement: all gpu tests, the new large-shape parity tests, the default (cfg4) bench line and the
# ncu launch list + GEMM DRAM bytes of one timed cfg4 step
out=gpurun_out
python -m pytest tests -m gpu -q -rA -x --deselect tests/test_gpu_parity_large.py 2>&1 | tail -15 > $out/r2a_pytest_old.log
python -m pytest tests/test_gpu_parity_large.py -m gpu -q -rA -s 2>&1 | tail -150 > $out/r2a_pytest_large.log
python bench.py --steps 10 --warmup 3 > $out/r2a_bench_cfg4.json 2> $out/r2a_bench_cfg4.err
tail -c 600 $out/r2a_bench_cfg4.err
export LAKO_PROFILER_RANGE=1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/r2a_launches_bench_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras \
    > $out/r2a_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:gemm2 --csv --log-file $out/r2a_gemm_dram_cfg4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $out/r2a_gemm_dram.log 2>&1
unset LAKO_PROFILER_RANGE
tail -5 $out/r2a_pytest_old.log
grep -E "passed|failed|PASS|FAIL|XFAIL|XPASS|Error" $out/r2a_pytest_large.log | tail -20
