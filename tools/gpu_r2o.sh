#!/bin/bash
out=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'enc_attn_bwd_kernel' --launch-skip 2 -c 1 -f -o /tmp/r2o_bwd \
    python tools/attn_bwd_time.py > /dev/null 2>&1
ncu -i /tmp/r2o_bwd.ncu-rep --page raw --csv > $out/r2o_ncu_full_attn_bwd.csv 2>/dev/null
ncu -i /tmp/r2o_bwd.ncu-rep --page source --csv > $out/r2o_ncu_source_attn_bwd.csv 2>/dev/null
ls -la $out/r2o*
