#!/bin/bash
# per-source-line profile of the transmitter-chain kernel (K=64 and K=1024 shapes), summarised on the box
OUT=gpurun_out
mkdir -p $OUT
for spec in "k64 0" "k1024 4"; do
  set -- $spec
  CHAIN_STEPS=1 CHAIN_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:fused_mod_kernel -s $2 -c 1 -f -o /tmp/tx_$1 python tools/chain_bench.py tx > $OUT/r01k_tx_$1_ncu.log 2>&1
  ncu -i /tmp/tx_$1.ncu-rep --page source --print-source cuda,sass --csv > /tmp/tx_$1_src.csv 2>/dev/null
  python tools/ncu_lines.py /tmp/tx_$1_src.csv 45 > $OUT/r01k_tx_$1_lines.txt 2>&1
  ncu -i /tmp/tx_$1.ncu-rep --page raw --csv > /tmp/tx_$1_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/tx_$1_raw.csv > $OUT/r01k_tx_$1_summary.txt 2>&1
done
ls -la $OUT
