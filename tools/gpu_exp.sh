#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/attn_bwd_time.py
