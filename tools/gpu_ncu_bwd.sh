#!/bin/bash
out=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'enc_attn_bwd_kernel' --launch-skip 2 -c 1 -f -o /tmp/bwd python tools/attn_bwd_time.py > /dev/null 2>&1
ncu -i /tmp/bwd.ncu-rep --page source --csv > $out/r2_ncu_source_attn_bwd.csv 2>/dev/null
ls -la $out/r2_ncu_source_attn_bwd.csv
