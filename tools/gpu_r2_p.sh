#!/bin/bash
# Round-2 GPU pass P: parity suite after the generic-kernel rewrite (composite radices), shape rows, ncu summary of the generic modulator.
TAG=${1:-r02p}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== rows" ; timeout 400 python tools/chain_bench.py shapes > $OUT/${TAG}_chain.jsonl 2> $OUT/${TAG}_chain.err; tail -n 3 $OUT/${TAG}_chain.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_smem_mod -s 3 -c 1 -f -o /tmp/gen_mod \
    python tools/chain_bench.py shapes > $OUT/${TAG}_ncu.log 2>&1
ncu -i /tmp/gen_mod.ncu-rep --page raw --csv > /tmp/gen_raw.csv 2>/dev/null && python tools/ncu_summary.py /tmp/gen_raw.csv > $OUT/${TAG}_generic_mod_ncu_full.txt
ncu -i /tmp/gen_mod.ncu-rep --page source --print-source cuda,sass --csv > /tmp/gen_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/gen_src.csv 40 > $OUT/${TAG}_generic_mod_lines.txt 2>&1
