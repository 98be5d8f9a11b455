#!/bin/bash
out=gpurun_out
python -m pytest tests/test_gpu_decode.py -m gpu -q -x 2>&1 | tail -25 > $out/r2b_pytest_decode.log
python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_base.py -m gpu -q -x 2>&1 | tail -25 > $out/r2b_pytest_model.log
python -m pytest tests/test_gpu_parity_large.py -m gpu -q -rA -s 2>&1 | tail -150 > $out/r2b_pytest_large.log
python bench.py --workload cfg3 --steps 5 --warmup 3 > $out/r2b_bench_cfg3.json 2> $out/r2b_bench_cfg3.err
tail -c 400 $out/r2b_bench_cfg3.err
python tools/e2e_probe.py cfg4 > $out/r2b_e2e_probe.log 2>&1
LAKO_DECODE_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:'xattn_decode_kernel|smallm_linear|dself_step|xattn_decode_combine' \
    --launch-skip 300 -c 24 -f -o /tmp/r2b_decode python tools/prof_decode.py > $out/r2b_ncu_decode.log 2>&1
ncu -i /tmp/r2b_decode.ncu-rep --page raw --csv > $out/r2b_ncu_full_decode.csv 2>/dev/null
LAKO_DECODE_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r2b_launches_decode.csv \
    python tools/prof_decode.py 8 20 250 3 > /dev/null 2>&1
tail -4 $out/r2b_pytest_decode.log; tail -4 $out/r2b_pytest_model.log
grep -E "passed|failed|PASS|FAIL|XFAIL|XPASS|Error" $out/r2b_pytest_large.log | tail -12
cat $out/r2b_e2e_probe.log
