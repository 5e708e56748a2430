#!/bin/bash
# Round-2 GPU pass N: parity suite after the generic-kernel twiddle change and the split transmitter store loop; shape and tx rows.
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== rows" ; timeout 400 python tools/chain_bench.py shapes tx > $OUT/${TAG}_chain.jsonl 2> $OUT/${TAG}_chain.err; tail -n 3 $OUT/${TAG}_chain.err
