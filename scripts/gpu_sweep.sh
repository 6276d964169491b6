#!/bin/bash
# parity smoke + knob sweeps of the persistent kernels.  usage: gpu_sweep.sh <tag> <n> <tune args...>
TAG=${1:-s}; shift
N=${1:-128}; shift
mkdir -p gpurun_out
LFMGPU_PIPE_DBG=0 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_fields_bit_exact_fp64 and (hex3d_m2_p4 or quad2d_m1 or hex3d_m1_p8)" > gpurun_out/${TAG}_first.log 2>&1
rc=$?; echo "rc=$rc"; tail -4 gpurun_out/${TAG}_first.log
[ $rc -ne 0 ] && exit $rc
LFMGPU_PLAN_STATS=1 timeout 600 python -m lfm_public_b200.tools.tune --n $N --steps 5 "$@" > gpurun_out/${TAG}_tune${N}.log 2>&1; echo "rc=$?"; grep -v "lfmgpu plan" gpurun_out/${TAG}_tune${N}.log | cut -c1-230
