#!/bin/bash
# round-2 refresh: all gpu tests, smoke(), default (cfg4) bench line with extras, reference arm, eval workloads,
# ncu launch list of one timed cfg4 step, ncu --set full of the attention kernels
out=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rA -x 2>&1 | tail -40 > $out/r2x_pytest_all.log
grep -E "passed|failed|xfailed|XPASS|XFAIL|FAILED" $out/r2x_pytest_all.log | tail -6
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1200 python bench.py --steps 10 --warmup 3 > $out/r2x_bench_cfg4.json 2> $out/r2x_bench_cfg4.err
head -c 400 $out/r2x_bench_cfg4.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/r2x_bench_reference.json 2> $out/r2x_bench_reference.err
head -c 400 $out/r2x_bench_reference.json; echo
timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > $out/r2x_bench_cfg3.json 2> $out/r2x_bench_cfg3.err
head -c 300 $out/r2x_bench_cfg3.json; echo
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > $out/r2x_bench_cfg5.json 2> $out/r2x_bench_cfg5.err
head -c 300 $out/r2x_bench_cfg5.json; echo
export LAKO_PROFILER_RANGE=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/r2x_launches_bench_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras \
    > $out/r2x_launches.log 2>&1
unset LAKO_PROFILER_RANGE
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'enc_attn_fwd2_kernel' --launch-skip 4 -c 1 -f -o /tmp/r2x_fwd \
    python tools/attn_fwd_trace.py > /dev/null 2>&1
ncu -i /tmp/r2x_fwd.ncu-rep --page raw --csv > $out/r2x_ncu_full_attn_fwd2.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'enc_attn_bwd_kernel' --launch-skip 2 -c 1 -f -o /tmp/r2x_bwd \
    python tools/attn_bwd_time.py > /dev/null 2>&1
ncu -i /tmp/r2x_bwd.ncu-rep --page raw --csv > $out/r2x_ncu_full_attn_bwd.csv 2>/dev/null
ls -la $out/r2x_*
