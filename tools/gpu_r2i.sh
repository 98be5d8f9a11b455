#!/bin/bash
out=gpurun_out
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "enc_attn" 2>&1 | tail -5 > $out/r2i_pytest_attn.log
tail -3 $out/r2i_pytest_attn.log
LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_trace.so timeout 300 python tools/attn_fwd_trace.py > $out/r2i_attn_trace.log 2>&1
cat $out/r2i_attn_trace.log
