#!/bin/bash
# Round 2, second session: cooperative producers (LFMGPU_PIPE_COOP), per-warp stores (LFMGPU_PIPE_WSTORE), L2 prefetch of the
# per-cell inputs of phase C (LFMGPU_PIPE_PFC).  Every step under its own timeout.   usage: gpu_r2b.sh <tag>
TAG=${1:-r4a}
mkdir -p gpurun_out
GOOD=""
for cfg in "1 1" "1 0" "0 1"; do
  set -- $cfg
  echo "=== COOP=$1 WSTORE=$2: one small case + one 2D case"
  LFMGPU_PIPE_COOP=$1 LFMGPU_PIPE_WSTORE=$2 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fields_bit_exact_fp64 and (hex3d_m2_p4 or quad2d_m1)" > gpurun_out/${TAG}_first_$1$2.log 2>&1
  rc=$?; echo "rc=$rc"; tail -4 gpurun_out/${TAG}_first_$1$2.log
  if [ $rc -eq 0 ] && [ -z "$GOOD" ]; then GOOD="$cfg"; fi
done
[ -z "$GOOD" ] && { echo "no new variant passes"; exit 1; }
set -- $GOOD
export LFMGPU_PIPE_COOP=$1 LFMGPU_PIPE_WSTORE=$2
echo "=== continuing with COOP=$1 WSTORE=$2"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact_fp64 or fp32 or medium" > gpurun_out/${TAG}_parity.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/${TAG}_parity.log
[ $rc -ne 0 ] && exit $rc
timeout 200 python -m pytest tests/test_zz_large.py -m gpu -x -q > gpurun_out/${TAG}_large.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/${TAG}_large.log
[ $rc -ne 0 ] && exit $rc
unset LFMGPU_PIPE_COOP LFMGPU_PIPE_WSTORE
LFMGPU_PLAN_STATS=1 timeout 240 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_PIPE_COOP=1,0 --set LFMGPU_PIPE_WSTORE=1,0 --set LFMGPU_PIPE_PFC=1,0 > gpurun_out/${TAG}_tune128.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_PIPE_PF=2,3,5 --set LFMGPU_PIPE_DBG=0,4 > gpurun_out/${TAG}_tune128_pf.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128_pf.log
LFMGPU_PLAN_STATS=1 timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_TILE_CELLS=64,96 > gpurun_out/${TAG}_tune128_tc.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*\|pipe\] stage.*' gpurun_out/${TAG}_tune128_tc.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --precision 4 --set LFMGPU_PIPE_COOP=1,0 > gpurun_out/${TAG}_tune128_fp32.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128_fp32.log
timeout 300 python -m lfm_public_b200.tools.tune --n 256 --steps 5 --tile morton --set LFMGPU_PIPE_COOP=1,0 > gpurun_out/${TAG}_tune256.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune256.log
