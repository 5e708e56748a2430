#!/bin/bash
# ncu --set full capture of the row kernels (transmitter chain, estimator, decision epilogue, burst extraction);
# summarised on the box (the report itself is too large to travel)
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python tools/chain_bench.py tx rx copy next > $OUT/${TAG}_chain_bench.jsonl 2> $OUT/${TAG}_chain_bench.err
CHAIN_STEPS=1 CHAIN_WARMUP=1 timeout 900 ncu --set full --clock-control none \
    -k regex:'fused_mod_kernel|est_fused|extract_burst|fused_rx_kernel' -c 36 -f -o /tmp/${TAG}_rows \
    python tools/chain_bench.py tx rx next > $OUT/${TAG}_rows_ncu.log 2>&1
ncu -i /tmp/${TAG}_rows.ncu-rep --page raw --csv > /tmp/${TAG}_rows_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/${TAG}_rows_raw.csv > $OUT/${TAG}_rows_summary.txt
ls -la $OUT /tmp/${TAG}_rows* | tail -8
