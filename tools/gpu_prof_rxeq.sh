#!/bin/bash
# per-source-line profile of the equalising receiver kernel at K=1024 (chain_bench rx: 2nd shape), summarised on the box
OUT=gpurun_out
mkdir -p $OUT
CHAIN_STEPS=1 CHAIN_WARMUP=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:fused_rx_kernel -s 6 -c 1 -f -o /tmp/rxeq python tools/chain_bench.py rx > $OUT/r01n_rxeq_ncu.log 2>&1
ncu -i /tmp/rxeq.ncu-rep --page source --print-source cuda,sass --csv > /tmp/rxeq_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/rxeq_src.csv 40 > $OUT/r01n_rxeq_lines.txt 2>&1
ncu -i /tmp/rxeq.ncu-rep --page raw --csv > /tmp/rxeq_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/rxeq_raw.csv > $OUT/r01n_rxeq_summary.txt 2>&1
