#!/bin/bash
out=gpurun_out
timeout 400 python bench.py --gpus 1 --steps 6 --warmup 3 --no-extras --no-cpu-baseline > $out/r2s_bench_cfg4_n1.json 2> $out/r2s_bench_cfg4_n1.err
head -c 330 $out/r2s_bench_cfg4_n1.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 6 --warmup 3 --no-extras --no-cpu-baseline > $out/r2s_bench_cfg4_n8.json 2> $out/r2s_bench_cfg4_n8.err
head -c 330 $out/r2s_bench_cfg4_n8.json; echo; tail -2 $out/r2s_bench_cfg4_n8.err
