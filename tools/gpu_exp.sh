#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_dropout.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/attn_bwd_time.py
timeout 600 python tools/dropout_cost.py 2>&1 | tail -2 | tee gpurun_out/r2p_dropout_cost.log
