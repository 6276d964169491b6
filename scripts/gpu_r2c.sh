#!/bin/bash
# Round 2, second session, call 2: gradient kernel with four consumer groups (LFMGPU_PIPE_GGROUPS) and per-warp stores
# (LFMGPU_PIPE_GWSTORE), stage kernel with an L1 prefetch of the second round's face constants (LFMGPU_PIPE_PFF).   usage: gpu_r2c.sh <tag>
TAG=${1:-r4b}
mkdir -p gpurun_out
GOOD=""
for cfg in "4 1" "4 0" "3 1"; do
  set -- $cfg
  echo "=== GGROUPS=$1 GWSTORE=$2 PFF=1: small cases"
  LFMGPU_PIPE_GGROUPS=$1 LFMGPU_PIPE_GWSTORE=$2 LFMGPU_PIPE_PFF=1 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fields_bit_exact_fp64 and (hex3d_m2_p4 or quad2d_m1)" > gpurun_out/${TAG}_first_$1$2.log 2>&1
  rc=$?; echo "rc=$rc"; tail -4 gpurun_out/${TAG}_first_$1$2.log
  if [ $rc -eq 0 ] && [ -z "$GOOD" ]; then GOOD="$cfg"; fi
done
[ -z "$GOOD" ] && { echo "no new variant passes"; exit 1; }
set -- $GOOD
export LFMGPU_PIPE_GGROUPS=$1 LFMGPU_PIPE_GWSTORE=$2 LFMGPU_PIPE_PFF=1
echo "=== continuing with GGROUPS=$1 GWSTORE=$2 PFF=1"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact_fp64 or fp32 or medium" > gpurun_out/${TAG}_parity.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/${TAG}_parity.log
[ $rc -ne 0 ] && exit $rc
timeout 200 python -m pytest tests/test_zz_large.py -m gpu -x -q > gpurun_out/${TAG}_large.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/${TAG}_large.log
[ $rc -ne 0 ] && exit $rc
unset LFMGPU_PIPE_GGROUPS LFMGPU_PIPE_GWSTORE LFMGPU_PIPE_PFF
LFMGPU_PLAN_STATS=1 timeout 240 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_PIPE_GGROUPS=4,3 --set LFMGPU_PIPE_GWSTORE=1,0 > gpurun_out/${TAG}_tune128.log 2>&1; echo "rc=$?"; grep -o 'pipe\] stage.*\|"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_PIPE_PFF=0,1 --set LFMGPU_PIPE_PF=2,3 > gpurun_out/${TAG}_tune128_pf.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128_pf.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_PIPE_GPF=0,2,3 --set LFMGPU_PIPE_GSLOTS=12,4 > gpurun_out/${TAG}_tune128_gpf.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128_gpf.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --precision 4 --set LFMGPU_PIPE_GGROUPS=4,3 --set LFMGPU_PIPE_PFF=0,1 > gpurun_out/${TAG}_tune128_fp32.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128_fp32.log
timeout 300 python -m lfm_public_b200.tools.tune --n 256 --steps 5 --tile morton --set LFMGPU_PIPE_GGROUPS=4,3 --set LFMGPU_PIPE_PFF=0,1 > gpurun_out/${TAG}_tune256.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune256.log
