#!/bin/bash
# round 2, first GPU call: baseline timing, the EARLY copy-in variant, source-level stall sampling of k_tile_stage
set -x
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
LFMGPU_STAGE_CFG=7 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bit_exact_fp64 or medium" > gpurun_out/${TAG}_early_parity.log 2>&1; tail -2 gpurun_out/${TAG}_early_parity.log
timeout 300 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --set LFMGPU_STAGE_CFG=6,7 > gpurun_out/${TAG}_early_tune.log 2>&1; tail -2 gpurun_out/${TAG}_early_tune.log
timeout 300 python -m lfm_public_b200.tools.tune --n 256 --steps 5 --set LFMGPU_STAGE_CFG=6,7 > gpurun_out/${TAG}_early_tune256.log 2>&1; tail -2 gpurun_out/${TAG}_early_tune256.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 10 -c 2 -o gpurun_out/${TAG}_stage128 \
	python -m lfm_public_b200.tools.tune --n 128 --steps 1 > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
