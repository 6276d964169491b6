#!/bin/bash
# First GPU call of the next round: the checks written after round 1's GPU budget was spent, then the
# measurements that decide which stage-kernel experiment to run first.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first_call.sh r2a'
set -x
TAG=${1:-r2a}
mkdir -p gpurun_out
# 1. full-size parity (256^3, one time step against the serial oracle, both kernel modes) + the 128^3 case of the suite
LFM_FULL_SIZE_TESTS=1 timeout 900 python -m pytest tests/test_zz_large.py -m gpu -q > gpurun_out/${TAG}_full_size_parity.log 2>&1; tail -3 gpurun_out/${TAG}_full_size_parity.log
# 2. solver 2 in fp32 (is the AUSM stage still arithmetic-bound there?) next to fp64
for P in 8 4; do timeout 120 python -m lfm_public_b200.tools.tune --n 128 --scheme 2 --minmod --precision $P --steps 3; done > gpurun_out/${TAG}_ausm_fp64_fp32.log 2>&1; tail -2 gpurun_out/${TAG}_ausm_fp64_fp32.log
# 3. the EARLY copy-in variant (DESIGN.md 9 a'): parity first, then its time next to the default
LFMGPU_STAGE_CFG=7 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bit_exact_fp64 or medium" > gpurun_out/${TAG}_early_parity.log 2>&1; tail -2 gpurun_out/${TAG}_early_parity.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --set LFMGPU_STAGE_CFG=6,7 > gpurun_out/${TAG}_early_tune.log 2>&1; tail -2 gpurun_out/${TAG}_early_tune.log
# 4. where the stage kernel waits: full ncu capture with source-level stall sampling of 4 launches at 128^3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_stage -s 10 -c 4 -o gpurun_out/${TAG}_stage128 \
	python -m lfm_public_b200.tools.tune --n 128 --steps 1 > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
# read here with:
#   ncu -i gpurun_out/${TAG}_stage128.ncu-rep --page raw --csv | grep -E "smsp__average_warp.*stall|l1tex__t_(requests|sectors).*ldgsts|sm__warps_active|dram__bytes"
#   ncu -i gpurun_out/${TAG}_stage128.ncu-rep --page source --csv   (stall samples per source line: the wait_group, the barriers, fetch_face)
