#!/bin/bash
# compute-sanitizer over one small invocation of every fused kernel.  Usage: bash tools/gpu_sanitize.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  for k in 16 64 256 1024 2048 128; do
    echo "== compute-sanitizer --tool $tool, K=$k"
    timeout 300 compute-sanitizer --tool $tool --print-limit 10 --log-file $OUT/${TAG}_sanitizer_${tool}_k$k.log \
        python tools/sanitize_target.py $k > $OUT/${TAG}_sanitizer_${tool}_k${k}_stdout.txt 2>&1
    echo "rc=$?" >> $OUT/${TAG}_sanitizer_${tool}_k${k}_stdout.txt
    tail -2 $OUT/${TAG}_sanitizer_${tool}_k$k.log
    tail -2 $OUT/${TAG}_sanitizer_${tool}_k${k}_stdout.txt
  done
done
