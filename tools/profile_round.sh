#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box, one GPU):  tools/profile_round.sh <tag>
#  1. launch list of ONE timed training step of the bench command (per-launch durations, cold cache, serialised)
#  2. DRAM bytes of every GEMM launch of that step (roofline.traffic)
#  3. --set full of the hot kernels at the cfg2 shapes, exported as raw CSV (the .ncu-rep stays on the box)
tag=${1:-r1}
out=gpurun_out
export LAKO_PROFILER_RANGE=1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches_bench_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
    > $out/${tag}_launches.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:gemm --csv --log-file $out/${tag}_gemm_dram_cfg2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/${tag}_gemm_dram.log 2>&1
unset LAKO_PROFILER_RANGE
ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16|enc_attn|rmsnorm' -c 20 -f -o /tmp/${tag}_full \
    python tools/prof_kernels.py > $out/${tag}_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_ncu_full_hot_kernels.csv 2>/dev/null
ls -la $out/${tag}_*
