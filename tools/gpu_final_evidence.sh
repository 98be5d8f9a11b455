#!/bin/bash
# final-code ncu evidence: launch list of one timed cfg4 step, launch list of a greedy decode (graph off so that every
# launch is visible), --set full of the encoder attention backward
out=gpurun_out
export LAKO_PROFILER_RANGE=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/r2f_launches_bench_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras \
    > $out/r2f_launches.log 2>&1
unset LAKO_PROFILER_RANGE
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2f_launches_decode_step.csv \
    python tools/prof_decode.py 8 20 250 4 > $out/r2f_decode.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'enc_attn_bwd_kernel' --launch-skip 2 -c 1 -f -o /tmp/r2f_bwd \
    python tools/attn_bwd_time.py > /dev/null 2>&1
ncu -i /tmp/r2f_bwd.ncu-rep --page raw --csv > $out/r2f_ncu_full_attn_bwd.csv 2>/dev/null
ls -la $out/r2f_*
