#!/bin/bash
# Round-2 GPU pass R: per-stage cycles of the two-pass kernels (parked modulator and the L2-scratch version).
TAG=${1:-r02r}
mkdir -p gpurun_out
timeout 200 python tools/stage_profile.py c5 2048 2>&1 | tee gpurun_out/${TAG}_stage_cycles_c5_parked.txt
GFDM_MOD2_SCRATCH=1 timeout 200 python tools/stage_profile.py c5 2048 2>&1 | head -8 | tee gpurun_out/${TAG}_stage_cycles_c5_scratch.txt
