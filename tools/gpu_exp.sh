#!/bin/bash
for f in lako_b200/_lib/liblako_b200_paced.so lako_b200/_lib/liblako_b200.so lako_b200/_lib/liblako_b200_paced.so lako_b200/_lib/liblako_b200.so; do echo "== in-step $f"
LAKO_B200_LIB=$PWD/$f timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | cut -c1-100
done
LAKO_B200_LIB=$PWD/lako_b200/_lib/liblako_b200_paced.so timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
