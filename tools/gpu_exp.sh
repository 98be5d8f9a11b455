#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_attention.py tests/test_gpu_dropout.py tests/test_gpu_parity_base.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | tee gpurun_out/r2t_dropout_cost.log
