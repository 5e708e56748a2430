#!/bin/bash
# compute-sanitizer over one small invocation of every fused kernel.  Usage: bash tools/gpu_sanitize.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 --log-file $OUT/${TAG}_sanitizer_$tool.log \
      python tools/sanitize_target.py > $OUT/${TAG}_sanitizer_${tool}_stdout.txt 2>&1
  echo "rc=$?" >> $OUT/${TAG}_sanitizer_${tool}_stdout.txt
  tail -5 $OUT/${TAG}_sanitizer_$tool.log
  tail -3 $OUT/${TAG}_sanitizer_${tool}_stdout.txt
done
