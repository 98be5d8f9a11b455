#!/bin/bash
# ncu --set full of the three any-length attention kernels (enc_attn_long.cu) at P=80 L=512 H=16
out=gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:'enc_attn_long_fwd' --launch-skip 4 -c 1 -f -o /tmp/long_f python tools/attn_bwd_time.py 80 512 16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'enc_attn_long_d' --launch-skip 4 -c 2 -f -o /tmp/long_b python tools/attn_bwd_time.py 80 512 16 > /dev/null 2>&1
ncu -i /tmp/long_f.ncu-rep --page raw --csv > $out/r2_ncu_full_enc_attn_long.csv 2>/dev/null
ncu -i /tmp/long_b.ncu-rep --page raw --csv | tail -n +3 >> $out/r2_ncu_full_enc_attn_long.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r2_ncu_full_enc_attn_long.csv")))
h=rows[0]
want=["Kernel Name","gpu__time_duration.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__inst_executed_pipe_tensor.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","smsp__issue_active.avg.pct","sm__inst_executed_pipe_xu.sum"]
idx=[h.index(w) for w in want if w in h]
for r in rows[2:]:
    print([ (h[i], r[i]) for i in idx])
PY
