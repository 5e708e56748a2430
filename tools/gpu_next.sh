#!/bin/bash
# GPU pass for the rows either side of the path (tests/test_next_rows.py) + the headline bench with the chunk chain.
TAG=${1:-r01e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest next rows" ; timeout 900 python -m pytest tests/test_next_rows.py -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_next.txt
echo "== topo" ; (nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^CPU(s)"; nproc) > $OUT/${TAG}_topo.txt 2>&1
echo "== bench c3" ; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee $OUT/${TAG}_bench_c3.json
