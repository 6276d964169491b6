set -x
mkdir -p gpurun_out
timeout 600 python -m lfm_public_b200.tools.tune --n 128 --steps 3 --set LFMGPU_STAGE_CFG=6,7 > gpurun_out/r21.log 2>&1; cat gpurun_out/r21.log
