#!/bin/bash
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp_nccl.py -m gpu -q -x 2>&1 | tail -15 > $out/r2q_pytest_dp_nccl.log
cat $out/r2q_pytest_dp_nccl.log
timeout 300 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3
