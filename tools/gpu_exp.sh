#!/bin/bash
LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_exp1.so timeout 300 python tools/attn_fwd_trace.py 2>&1 | head -3
LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_exp1.so timeout 300 python tools/attn_fwd_trace.py 2>&1 | grep -A6 "softmax group 0"
