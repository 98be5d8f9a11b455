#!/bin/bash
# same-box A/B of two library builds (expects _lib/liblako_b200_old.so next to the new one)
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_dropout.py -m gpu -q -x 2>&1 | tail -2
for f in old "" old "" old ""; do
lib=lako_b200/_lib/liblako_b200${f:+_$f}.so
echo "== $lib"
LAKO_B200_LIB=$PWD/$lib timeout 300 python tools/attn_bwd_time.py | tail -1
LAKO_B200_LIB=$PWD/$lib timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | head -1 | cut -c1-100
done
