#!/bin/bash
out=gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -5 > $out/r2l_pytest.log
tail -3 $out/r2l_pytest.log
timeout 300 python tools/attn_fwd_trace.py 2>&1 | head -2
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-extras > $out/r2l_bench_cfg4.json 2> $out/r2l_bench_cfg4.err
cat $out/r2l_bench_cfg4.json | head -c 1500
LAKO_PROF=1 timeout 600 python bench.py --workload cfg4 --steps 4 --warmup 3 --no-extras 2>&1 | grep -v "^{" | head -40 > $out/r2l_prof_cfg4.log
