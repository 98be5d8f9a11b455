#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_dropout.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | tee gpurun_out/r2n_dropout_cost.log
