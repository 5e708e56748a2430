#!/bin/bash
# per-source-line profile (ncu --set full --import-source on) of the two headline kernels at C3, summarised on the box
OUT=gpurun_out
mkdir -p $OUT
for k in fused_mod_kernel fused_rx_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o /tmp/c3_$k \
      python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/r01x_c3_${k}_ncu.log 2>&1
  ncu -i /tmp/c3_$k.ncu-rep --page source --print-source cuda,sass --csv > /tmp/c3_${k}_src.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/c3_${k}_src.csv 45 > $OUT/r01x_c3_${k}_lines.txt 2>&1
done
ls -la $OUT | tail -6
