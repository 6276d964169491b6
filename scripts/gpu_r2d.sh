#!/bin/bash
# Round 2, second session, call 3: finished records written from registers straight to HBM (LFMGPU_PIPE_DIRECT, bit 0 stage kernel,
# bit 1 gradient kernel).   usage: gpu_r2d.sh <tag>
TAG=${1:-r4c}
mkdir -p gpurun_out
GOOD=""
for cfg in 3 2 1; do
  echo "=== DIRECT=$cfg: small cases"
  LFMGPU_PIPE_DIRECT=$cfg timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fields_bit_exact_fp64 and (hex3d_m2_p4 or quad2d_m1)" > gpurun_out/${TAG}_first_$cfg.log 2>&1
  rc=$?; echo "rc=$rc"; tail -4 gpurun_out/${TAG}_first_$cfg.log
  if [ $rc -eq 0 ] && [ -z "$GOOD" ]; then GOOD="$cfg"; fi
done
[ -z "$GOOD" ] && { echo "no new variant passes"; exit 1; }
export LFMGPU_PIPE_DIRECT=$GOOD
echo "=== continuing with DIRECT=$GOOD"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact_fp64 or fp32 or medium" > gpurun_out/${TAG}_parity.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/${TAG}_parity.log
[ $rc -ne 0 ] && exit $rc
timeout 200 python -m pytest tests/test_zz_large.py -m gpu -x -q > gpurun_out/${TAG}_large.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/${TAG}_large.log
[ $rc -ne 0 ] && exit $rc
unset LFMGPU_PIPE_DIRECT
LFMGPU_PLAN_STATS=1 timeout 240 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_PIPE_DIRECT=0,1,2,3 --set LFMGPU_PIPE_GGROUPS=3,4 > gpurun_out/${TAG}_tune128.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --set LFMGPU_PIPE_DIRECT=3 --set LFMGPU_PIPE_GPF=0,2 --set LFMGPU_PIPE_DBG=0,4 > gpurun_out/${TAG}_tune128_gpf.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128_gpf.log
timeout 200 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --tile morton --precision 4 --set LFMGPU_PIPE_DIRECT=0,2,3 > gpurun_out/${TAG}_tune128_fp32.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune128_fp32.log
timeout 300 python -m lfm_public_b200.tools.tune --n 256 --steps 5 --tile morton --set LFMGPU_PIPE_DIRECT=0,2,3 > gpurun_out/${TAG}_tune256.log 2>&1; echo "rc=$?"; grep -o '"knobs.*Gcell_stages_per_s": [0-9.]*' gpurun_out/${TAG}_tune256.log
