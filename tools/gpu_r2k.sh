#!/bin/bash
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "enc_attn" 2>&1 | tail -5 > $out/r2k_pytest_attn.log
tail -2 $out/r2k_pytest_attn.log
LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_trace.so timeout 300 python tools/attn_fwd_trace.py > $out/r2k_attn_trace.log 2>&1
head -3 $out/r2k_attn_trace.log; grep -A12 "softmax group 0" $out/r2k_attn_trace.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'enc_attn_fwd2_kernel' --launch-skip 4 -c 1 -f -o /tmp/r2k_attn \
    python tools/attn_fwd_trace.py > /dev/null 2>&1
ncu -i /tmp/r2k_attn.ncu-rep --page raw --csv > $out/r2k_ncu_full_attn_fwd2.csv 2>/dev/null
ncu -i /tmp/r2k_attn.ncu-rep --page source --csv > $out/r2k_ncu_source_attn_fwd2.csv 2>/dev/null
