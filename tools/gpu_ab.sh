#!/bin/bash
# same-box A/B of two library builds on the per-category step times (expects _lib/liblako_b200_old.so next to the new one)
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_dropout.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
for f in old "" old "" old ""; do
lib=lako_b200/_lib/liblako_b200${f:+_$f}.so
echo "== $lib"
LAKO_B200_LIB=$PWD/$lib timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | head -1 | cut -c1-100
done
