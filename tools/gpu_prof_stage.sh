#!/bin/bash
# per-stage cycle counters of an instrumented build (make -C gr-gfdm_b200 prof), production bench line beside it
TAG=${1:-r02i}
mkdir -p gpurun_out
GFDM_PROF_LIB=$PWD/gr-gfdm_b200/lib/libgfdm_b200_prof_nofence.so timeout 200 python tools/stage_profile.py c3 4096 2>&1 | tee gpurun_out/${TAG}_stage_cycles_c3_nofence.txt | head -12
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 | tee gpurun_out/${TAG}_bench_c3.json | cut -c1-200
