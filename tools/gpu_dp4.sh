#!/bin/bash
out=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 6 --warmup 3 --no-extras --no-cpu-baseline > $out/r2y_bench_cfg4_n4.json 2> $out/r2y_bench_cfg4_n4.err
head -c 330 $out/r2y_bench_cfg4_n4.json; echo; tail -2 $out/r2y_bench_cfg4_n4.err
