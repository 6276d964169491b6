#!/bin/bash
# staged first runs of the persistent kernels, each under its own short timeout (a deadlocked kernel must not eat the call);
# stops at the first stage that fails or hangs.   usage: gpu_pipe_check.sh <tag> [pipe values, default "3"]
TAG=${1:-p}
PIPES=${2:-3}
mkdir -p gpurun_out
for P in $PIPES; do
  echo "=== LFMGPU_PIPE=$P: one small case"
  LFMGPU_PIPE=$P timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fields_bit_exact_fp64 and hex3d_m2_p4" > gpurun_out/${TAG}_pipe${P}_first.log 2>&1
  rc=$?; echo "rc=$rc"; tail -12 gpurun_out/${TAG}_pipe${P}_first.log
  [ $rc -ne 0 ] && exit $rc
  echo "=== LFMGPU_PIPE=$P: zoo parity"
  LFMGPU_PIPE=$P timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bit_exact_fp64 or fp32 or medium" > gpurun_out/${TAG}_pipe${P}_parity.log 2>&1
  rc=$?; echo "rc=$rc"; tail -12 gpurun_out/${TAG}_pipe${P}_parity.log
  [ $rc -ne 0 ] && exit $rc
done
echo "=== 128^3 one step vs the oracle"
timeout 200 python -m pytest tests/test_zz_large.py -m gpu -x -q > gpurun_out/${TAG}_large.log 2>&1; rc=$?; echo "rc=$rc"; tail -6 gpurun_out/${TAG}_large.log
[ $rc -ne 0 ] && exit $rc
timeout 150 python -m lfm_public_b200.tools.tune --n 128 --steps 5 --set LFMGPU_PIPE=3,1,2,0 > gpurun_out/${TAG}_tune128.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/${TAG}_tune128.log | cut -c1-330
[ $rc -ne 0 ] && exit $rc
timeout 300 python -m lfm_public_b200.tools.tune --n 256 --steps 5 --set LFMGPU_PIPE=3,0 > gpurun_out/${TAG}_tune256.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/${TAG}_tune256.log | cut -c1-330
