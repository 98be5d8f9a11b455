#!/bin/bash
for f in lako_b200/_lib/liblako_b200.so lako_b200/_lib/liblako_b200_old.so lako_b200/_lib/liblako_b200.so lako_b200/_lib/liblako_b200_old.so; do echo "== $f"
LAKO_B200_LIB=$PWD/$f timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | head -1 | cut -c1-135
done
timeout 600 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_dropout.py -m gpu -q -x 2>&1 | tail -2
