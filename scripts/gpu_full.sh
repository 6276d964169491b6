#!/bin/bash
# full GPU suite + 256^3 timings.  usage: gpu_full.sh <tag>
TAG=${1:-f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
LFMGPU_PLAN_STATS=1 timeout 400 python -m lfm_public_b200.tools.tune --n 256 --steps 5 --set LFMGPU_PIPE_PF=0,3 > gpurun_out/${TAG}_tune256.log 2>&1; echo "rc=$?"; grep -v "lfmgpu plan" gpurun_out/${TAG}_tune256.log | cut -c1-250
LFMGPU_PLAN_STATS=1 timeout 400 python -m lfm_public_b200.tools.tune --n 256 --steps 5 --tile morton > gpurun_out/${TAG}_tune256_morton.log 2>&1; echo "rc=$?"; grep -v "lfmgpu plan" gpurun_out/${TAG}_tune256_morton.log | cut -c1-250
